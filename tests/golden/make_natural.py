#!/usr/bin/env python
"""Natural-image golden vectors (build container only: needs /root/reference and its images/training_set/).

LR / HR pairs are made the way the reference's dataset does (data_handling/data_class.py:61-68: PIL image ->
transforms.Resize(size) -> ToTensor) from a crop of one training image; the LR frame is quantised to uint8 so that it is
stored exactly.  The UNMODIFIED reference modules (fp32, CPU, random-init weights from oracle.weights) produce the
reference output; the tests check max-abs / PSNR against it and the PSNR DELTA against the HR target
(PSNR(ours, HR) - PSNR(reference, HR), BASELINE.json north_star: <= 0.05 dB).

  python tests/golden/make_natural.py        -> tests/golden/natural_<case>.npz  {lr_u8, hr_u8, ref (fp16)}
"""
import os, sys
import numpy as np
import torch
from PIL import Image
from torchvision import transforms

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, "/root/reference")
from oracle.weights import synth_state_dict      # noqa: E402
from tests.golden._refload import ref_model     # noqa: E402

NATURAL = {
    # name: (model, weight seed, image, crop box (left, top, right, bottom) of the 3840x2160 frame, LR size, HR size, forward kwargs)
    "natural_window_96x176_r1p5": ("WindowTransformer", 31, "image_103.png", (1200, 600, 1200 + 1056, 600 + 576), (96, 176), (144, 264),
                                   dict(res_out=(144, 264))),
    "natural_fast_96x176_x2": ("FastTransformer", 32, "image_104.png", (800, 400, 800 + 1056, 400 + 576), (96, 176), (192, 352),
                               dict(upscale_factor=2)),
}


# whole 720p frames: LR = the 4K training image resized to 720p exactly as data_class.py:61-64 does, HR = the
# same image resized to the model's output size.  The LR frame is stored whole (uint8); HR and the reference output are stored on
# a stride-3 lattice of the output (1/9 of the pixels: the PSNRs of the test are taken over that lattice for both sides).
NATURAL_FULL = {
    "natural_window_720p_1080p": ("WindowTransformer", 33, "image_103.png", (720, 1280), (1080, 1920), dict(res_out=(1080, 1920))),
    "natural_fast_720p_x2": ("FastTransformer", 34, "image_104.png", (720, 1280), (1440, 2560), dict(upscale_factor=2)),
    "natural_residual_720p_1080p": ("ResidualTransformer", 35, "image_109.png", (720, 1280), (1080, 1920), dict(res_out=(1080, 1920))),
}


def main_full():
    torch.set_num_threads(os.cpu_count())
    for name, (model, wseed, img, lr_size, hr_size, kw) in NATURAL_FULL.items():
        im = Image.open(os.path.join("/root/reference/images/training_set", img)).convert("RGB")
        lr = transforms.Compose([transforms.Resize(lr_size), transforms.ToTensor()])(im)
        hr = transforms.Compose([transforms.Resize(hr_size), transforms.ToTensor()])(im)
        lr_u8 = (lr * 255).round().clamp(0, 255).to(torch.uint8)
        hr_u8 = (hr * 255).round().clamp(0, 255).to(torch.uint8)
        x = (lr_u8.float() / 255.0).unsqueeze(0)
        M = ref_model(model)
        M.load_state_dict(synth_state_dict(model, wseed), strict=True)
        with torch.no_grad():
            ref = M(x, **kw)
        assert tuple(ref.shape[2:]) == hr_size
        np.savez_compressed(os.path.join(HERE, name + ".npz"), lr_u8=lr_u8.numpy(), hr_u8=hr_u8.numpy()[:, ::3, ::3],
                            ref=ref[0].numpy()[:, ::3, ::3].astype(np.float16), shape=np.array(ref.shape))
        mse = ((ref[0] - hr_u8.float() / 255) ** 2).mean().item()
        print(name, tuple(ref.shape), "PSNR(reference, HR) %.3f dB" % (10 * np.log10(1.0 / mse)), "mean", ref.mean().item(), flush=True)


def main():
    if "--full" in sys.argv:
        return main_full()
    torch.set_num_threads(os.cpu_count())
    for name, (model, wseed, img, box, lr_size, hr_size, kw) in NATURAL.items():
        im = Image.open(os.path.join("/root/reference/images/training_set", img)).convert("RGB").crop(box)
        lr = transforms.Compose([transforms.Resize(lr_size), transforms.ToTensor()])(im)
        hr = transforms.Compose([transforms.Resize(hr_size), transforms.ToTensor()])(im)
        lr_u8 = (lr * 255).round().clamp(0, 255).to(torch.uint8)
        hr_u8 = (hr * 255).round().clamp(0, 255).to(torch.uint8)
        x = (lr_u8.float() / 255.0).unsqueeze(0)
        M = ref_model(model)
        M.load_state_dict(synth_state_dict(model, wseed), strict=True)
        with torch.no_grad():
            ref = M(x, **kw)
        assert tuple(ref.shape[2:]) == hr_size
        np.savez_compressed(os.path.join(HERE, name + ".npz"), lr_u8=lr_u8.numpy(), hr_u8=hr_u8.numpy(),
                            ref=ref[0].numpy().astype(np.float16))
        mse = ((ref[0] - hr_u8.float() / 255) ** 2).mean().item()
        print(name, tuple(ref.shape), "PSNR(reference, HR) %.3f dB" % (10 * np.log10(1.0 / mse)), "mean", ref.mean().item())


if __name__ == "__main__":
    main()
