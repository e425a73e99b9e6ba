"""-m gpu: the ``FISRnet.build_model`` / ``train`` mirror (reference FISRnet.py:175-248, 580-743) on a tiny synthetic
training set written in the reference's on-disk formats (.npy stand-ins for the v7.3 .mat files, 5-D .flo)."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _write_set(d, n=6, h=32, w=32, seed=0):
    rng = np.random.default_rng(seed)
    # read_mat_file layout: [N, N_seq, C, W, H] uint8 (utils.py:29-42)
    np.save(os.path.join(d, "lr.npy"), rng.integers(0, 256, (n, 5, 3, w, h), dtype=np.uint8))
    np.save(os.path.join(d, "hr.npy"), rng.integers(0, 256, (n, 7, 3, 2 * w, 2 * h), dtype=np.uint8))
    from fisr_b200.utils import write_flo_file_5dim
    write_flo_file_5dim(rng.standard_normal((n, 8, h, w, 2)).astype(np.float32) * 2, os.path.join(d, "flow.flo"))
    write_flo_file_5dim(rng.standard_normal((n, 4, h, w, 2)).astype(np.float32) * 2, os.path.join(d, "flow_ss2.flo"))
    # read_mat_file_warp .npy layout: [N, N_seq, H, W, C], values 0..255
    np.save(os.path.join(d, "warp.npy"), rng.uniform(0, 255, (n, 8, h, w, 3)).astype(np.float32))
    np.save(os.path.join(d, "warp_ss2.npy"), rng.uniform(0, 255, (n, 4, h, w, 3)).astype(np.float32))


def _args(d, **kw):
    a = dict(checkpoint_dir=os.path.join(d, "ckpt"), test_img_dir=os.path.join(d, "img"), text_dir=os.path.join(d, "txt"),
             log_dir=os.path.join(d, "log"), train_data_path=os.path.join(d, "lr.npy"), train_label_path=os.path.join(d, "hr.npy"),
             train_flow_data_path=os.path.join(d, "flow.flo"), train_flow_ss2_data_path=os.path.join(d, "flow_ss2.flo"),
             train_warped_data_path=os.path.join(d, "warp.npy"), train_wapred_ss2_data_path=os.path.join(d, "warp_ss2.npy"),
             exp_num=1, scale_factor=2, epoch=2, init_lr=1e-4, freq_display=1, lr_type="stair_decay",
             lr_stair_decay_points=[1, 2], lr_decreasing_factor=0.1, lr_linear_decay_point=1, batch_size=2, val_batch_size=2,
             val_data_size=2, n_train_img_showed=1, recn_lambda=1.0, tm1_lambda=1.0, tm2_lambda=0.1, tmm_lambda=1.0,
             td_lambda=0.1, ss2_lambda=1.0)
    a.update(kw)
    return SimpleNamespace(**a)


def test_train_two_epochs_checkpoint_and_resume(engine, tmp_path, capsys):
    from fisr_b200 import FISRnet
    d = str(tmp_path)
    _write_set(d)
    net = FISRnet(engine, _args(d))
    net.build_model()
    assert net.train_iter == 2 and net.val_iter == 1 and net.data.shape == (4, 32, 32, 15) and net.label.shape == (4, 64, 64, 21)
    w0 = engine.get_params() if net._initialized else None
    net.train()
    out = capsys.readouterr().out
    assert "# (average) Epoch: [   0], LR: 0.0001000000" in out and "# (average) Epoch: [   1]" in out
    assert out.count("######### Validation (average)") == 2
    assert net.global_step == 4
    # piecewise-constant schedule: boundaries at 1 and 2 epochs = steps 2 and 4 (x <= boundary keeps the earlier value)
    net.global_step = 2; assert net._lr(1) == pytest.approx(1e-4)
    net.global_step = 3; assert net._lr(1) == pytest.approx(1e-5)
    net.global_step = 5; assert net._lr(2) == pytest.approx(1e-6)
    ckpt = os.path.join(d, "ckpt", "FISRnet_exp1")
    # a TensorFlow V2 bundle under the Saver's names (FISRnet.py:1092-1099), with the optimizer state
    assert os.path.exists(os.path.join(ckpt, "FISRnet-4.index")) and os.path.exists(os.path.join(ckpt, "checkpoint"))
    from fisr_b200 import tf_checkpoint as T
    names = T.list_variables(os.path.join(ckpt, "FISRnet-4"))
    assert "FISRnet/level_1/enc/level_0/conv/0/w/Adam_1" in names and "beta1_power" in names and T.GLOBAL_STEP_NAME in names
    assert len(names) == 3 * 276 + 3
    trained = engine.get_params()
    # the weights moved, and a fresh object resumes from the checkpoint at step 4 with nothing left to do
    from fisr_b200.init import xavier_params
    init = xavier_params(seed=0, bias_std=0.0)
    k = "FISRnet/level_3/SR/conv/2/w"
    assert np.abs(trained[k] - np.asarray(init[k])).max() > 1e-6
    net2 = FISRnet(engine, _args(d))
    ok, step = net2.load(net2.checkpoint_dir)
    assert ok and step == 4
    assert all(np.array_equal(engine.get_params()[n], trained[n]) for n in trained)
    assert net2._adam_restored and engine.adam_steps == 4
    engine.set_precision("f16x3")


def _batch(B, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    data = torch.rand(B, h, w, 15, generator=g)
    flow = (torch.randn(B, h, w, 16, generator=g) * 4 / 96 / 2).clamp(-1, 1)
    flow2 = (torch.randn(B, h, w, 8, generator=g) * 8 / 96 / 2).clamp(-1, 1)
    warp = torch.rand(B, h, w, 24, generator=g)
    warp2 = torch.rand(B, h, w, 12, generator=g)
    label = torch.rand(B, 2 * h, 2 * w, 21, generator=g)
    return [t.cuda() for t in (data, flow, flow2, warp, warp2, label)]


def test_resume_is_seamless(engine, tmp_path):
    """Saver.restore brings back m, v and the beta powers (FISRnet.py:1101-1115): step 3 after a save / clobber / load must be
    bit-identical to step 3 of the uninterrupted run; restarting Adam from zero moments instead moves the weights elsewhere."""
    from fisr_b200 import FISRnet
    from fisr_b200.init import xavier_params
    engine.set_precision("f16x3")
    net = FISRnet(engine, _args(str(tmp_path)))
    engine.set_params(xavier_params(seed=5, bias_std=0.01))
    net._initialized = True
    engine.adam_reset(0)
    for t in (1, 2):
        engine.train_step(*_batch(1, 32, 32, 80 + t), lr=1e-4)
    net.save_checkpoint(net.checkpoint_dir, 2)
    engine.train_step(*_batch(1, 32, 32, 83), lr=1e-4)
    straight = engine.get_params()

    engine.set_params(xavier_params(seed=6, bias_std=0.0))            # clobber weights and optimizer state
    engine.adam_reset(0)
    ok, step = net.load(net.checkpoint_dir)
    assert ok and step == 2 and net._adam_restored and engine.adam_steps == 2
    engine.train_step(*_batch(1, 32, 32, 83), lr=1e-4)
    resumed = engine.get_params()
    assert all(np.array_equal(resumed[k], straight[k]) for k in straight)

    net.load(net.checkpoint_dir)
    engine.adam_reset(0)                                              # what round 1 did: zero moments after a resume
    engine.train_step(*_batch(1, 32, 32, 83), lr=1e-4)
    cold = engine.get_params()
    k = "FISRnet/level_3/dec/level_0/conv/0/w"
    assert np.abs(cold[k] - straight[k]).max() > 10 * np.abs(resumed[k] - straight[k]).max() + 1e-7
    engine.adam_reset(0)


def test_train_step_rescales_on_overflow(engine):
    """Dynamic loss scaling: an absurd scale overflows the fp16 gradient planes; fisr_train_backward reports it, fisr_train_step
    lowers the scale and still takes the step."""
    import fisr_b200
    from fisr_b200.init import xavier_params
    engine.set_precision("f16x3")
    engine.set_params(xavier_params(seed=7, bias_std=0.01))
    engine.adam_reset(0)
    batch = _batch(1, 32, 32, 90)
    try:
        engine.set_loss_scale(2.0 ** 40)
        with pytest.raises(fisr_b200.FisrError, match="non-finite gradient"):
            engine.train_backward(*batch)
        before = engine.get_params()["FISRnet/level_3/SR/conv/2/w"].copy()
        s = engine.train_step(*batch, lr=1e-4)
        assert np.isfinite(s["total_loss"]) and engine.adam_steps == 1
        after = engine.get_params()["FISRnet/level_3/SR/conv/2/w"]
        assert np.isfinite(after).all() and np.abs(after - before).max() > 1e-6
    finally:
        engine.set_loss_scale(0.0)
        engine.adam_reset(0)


def test_validate_matches_torch(engine, tmp_path):
    from fisr_b200 import FISRnet
    from oracle import fisrnet_oracle as O
    from oracle import pipeline_oracle as P
    d = str(tmp_path)
    _write_set(d, n=4)
    net = FISRnet(engine, _args(d))
    net.build_model()
    params = O.init_params(3)
    engine.set_precision("f16x3")
    engine.set_params(params)
    net._initialized = True
    recn, psnr = net.validate(net.data_val, net.label_val, net.flow_val, net.warp_val)
    data, flow, warp, label = (torch.from_numpy(a) for a in (net.data_val, net.flow_val, net.warp_val, net.label_val))
    preds = [P.split_seq_dim(O.model(params, P.window_input(data, flow, warp, i))[2]) for i in range(3)]
    seq = P.groups2ovlp(torch.cat(preds, 1))
    gt = P.split_seq_dim(label)
    err = (seq - gt) ** 2
    assert recn == pytest.approx(float(err.mean()), rel=1e-4)
    assert psnr == pytest.approx(float((-10 * torch.log10(err.mean(dim=(2, 3, 4)))).mean()), abs=1e-3)
