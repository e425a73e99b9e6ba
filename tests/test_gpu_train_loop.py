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
    assert os.path.exists(os.path.join(ckpt, "FISRnet-4.npz")) and os.path.exists(os.path.join(ckpt, "checkpoint"))
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
    engine.set_precision("f16x3")


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
