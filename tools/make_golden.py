"""Generates the committed fixtures under tests/golden/ (run in the build container, where /root/reference exists).

 * scene1_yuv_rgb.npz  -- 4 crops (256x256) of the reference's shipped outputs
                          FISR_test_folder/scene1/FISR_frames/pred_YUV_k.png and pred_k.png (k = 0, 3), plus the full-frame
                          sizes of the shipped inputs / outputs.  Pins YUV2RGB_matlab + uint8 truncation (FISRnet.py:1066-1070)
                          and the output geometry 2048x3840 for a 1080x1920 input with the (2,2) tile grid.
 * scene1_lr_crop.npz  -- a 128x192 crop of the 5 shipped LR frames (real YUV statistics for window/warp tests).
 * model_fp64_64x96.npz -- oracle fp64 outputs of FISRnet.model for seeded weights (seed 7) and input (seed 8) at
                          1x64x96: golden vector for the CUDA path and for oracle regression.
"""
import os, sys
import numpy as np
import torch
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import fisrnet_oracle as O

REF = "/root/reference/FISR_test_folder/scene1"
OUT = os.path.join(ROOT, "tests", "golden")
os.makedirs(OUT, exist_ok=True)

crops = {}
for k, (y, x) in ((0, (300, 500)), (0, (1500, 2900)), (3, (800, 1200)), (3, (40, 3500))):
    yuv = np.array(Image.open(f"{REF}/FISR_frames/pred_YUV_{k}.png"))
    rgb = np.array(Image.open(f"{REF}/FISR_frames/pred_{k}.png"))
    crops[f"yuv_{k}_{y}_{x}"] = yuv[y:y + 256, x:x + 256]
    crops[f"rgb_{k}_{y}_{x}"] = rgb[y:y + 256, x:x + 256]
crops["output_hw"] = np.array(yuv.shape[:2])
lr = [np.array(Image.open(f"{REF}/LR_vid_1_fr_07171_seq_{i}.png")) for i in (1, 3, 5, 7, 9)]
crops["input_hw"] = np.array(lr[0].shape[:2])
crops["n_inputs"] = np.array(5)
crops["n_outputs"] = np.array(len([f for f in os.listdir(f"{REF}/FISR_frames") if f.startswith("pred_YUV_")]))
np.savez_compressed(os.path.join(OUT, "scene1_yuv_rgb.npz"), **crops)
np.savez_compressed(os.path.join(OUT, "scene1_lr_crop.npz"), frames=np.stack([f[400:528, 600:792] for f in lr]))

p64 = O.init_params(7, torch.float64)
x = O.synthetic_input(1, 64, 96, 8)
o = O.model(p64, x)
chk = float(sum(v.double().abs().sum() for v in p64.values()))
np.savez_compressed(os.path.join(OUT, "model_fp64_64x96.npz"), pred_l1=o[0].numpy().astype(np.float32),
                    pred_l2=o[1].numpy().astype(np.float32), pred_l3=o[2].numpy().astype(np.float32),
                    param_abs_sum=np.array(chk), input_sum=np.array(float(x.double().sum())))
print("golden written:", {f: os.path.getsize(os.path.join(OUT, f)) for f in os.listdir(OUT)})
