#!/usr/bin/env python
"""Evaluation driver with the reference's entry point and flags
(``experiment_scripts/eval_realestate10k.py``):

    python experiment_scripts/eval_realestate10k.py --experiment_name vis --batch_size 1 --gpus 1 \
        --views 2 --checkpoint_path model.pth --synthetic 8

Per scene: ``get_z`` once, then the whole 256x256 target in ONE ``forward`` call — the reference
splits it into 9 ray chunks to bound activation memory (eval_realestate10k.py:144-159); here the
chunking happens inside the library.  With ``--gpus N`` the rays of each scene are sharded across
the ranks and the tiles all-gathered (the reference's ranks each render everything).  Prints the
reference's ``elapsed`` and ``mse, psnr, lpip, ssim`` lines (eval_realestate10k.py:199): SSIM is the
restatement of the skimage call in ``cross_attention_renderer_b200/metrics.py``; LPIPS is computed when the
``lpips`` package and its weights are importable and printed as ``nan`` otherwise.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _common as C                                                   # noqa: E402
from cross_attention_renderer_b200 import metrics                     # noqa: E402


def multigpu_train(gpu, opt):
    dev = C.init_distributed(gpu, opt.gpus, opt.master_port)
    model = C.build_model(opt, dev)
    if opt.checkpoint_path is not None:
        C.load_checkpoint(model, opt.checkpoint_path, allow_encoder_mismatch=opt.allow_encoder_mismatch)
    model = model.eval()
    model.pixel_val_to_cpu = False
    if not opt.synthetic:
        raise RuntimeError("RealEstate10k frames are not available in this environment: pass --synthetic N")
    n_scenes = opt.synthetic if not opt.max_steps else min(opt.synthetic, opt.max_steps)
    H = opt.sidelength
    mses, psnrs, lpips_list, ssims = [], [], [], []
    loss_fn_alex = None
    if metrics.lpips_available():
        import lpips
        loss_fn_alex = lpips.LPIPS(net="vgg").to(dev)                 # eval_realestate10k.py:107
    with torch.no_grad():
        for val_i in range(n_scenes):
            model_input, gt = C.synthetic_scene_batch(1, H, 10_000 + val_i, device=dev)
            z = model.get_z(model_input)
            torch.cuda.synchronize()
            start = time.time()
            if opt.gpus > 1:
                out = C.sharding.render_sharded(model, model_input, z)
            else:
                out = model(model_input, z=z, val=True)
            torch.cuda.synchronize()
            if gpu == 0:
                print("elapsed: ", time.time() - start)
            rgb = out["rgb"].view(H, H, 3)
            valid_mask = out["valid_mask"].view(H, H, 1)
            target = gt["rgb"].view(H, H, 3)
            mse, psnr = C.psnr_masked(rgb, target, valid_mask)
            mses.append(mse)
            psnrs.append(psnr)
            # images as the reference scores them: [0, 1], invalid rays grey on both sides (:177-178)
            rgb01 = ((rgb + 1) * 0.5) * valid_mask + 0.5 * (1 - valid_mask)
            tgt01 = ((target + 1) * 0.5) * valid_mask + 0.5 * (1 - valid_mask)
            if loss_fn_alex is not None:
                lpips_list.append(float(loss_fn_alex(((rgb01.permute(2, 0, 1) - 0.5) * 2)[None],
                                                     ((tgt01.permute(2, 0, 1) - 0.5) * 2)[None])))
            else:
                lpips_list.append(float("nan"))
            ssims.append(metrics.ssim(rgb01, tgt01))                 # :195
            if gpu == 0:
                print("mse, psnr, lpip, ssim", np.mean(mses), np.mean(psnrs), np.mean(lpips_list), np.mean(ssims), flush=True)
    if opt.gpus > 1:
        torch.distributed.destroy_process_group()
    return float(np.mean(psnrs))


def main(argv=None):
    opt = C.base_parser(__doc__).parse_args(argv)
    C.spawn(multigpu_train, opt)


if __name__ == "__main__":
    main()
