#!/usr/bin/env python
"""Training driver with the reference's entry point and flags
(``experiment_scripts/train_realestate10k.py``, ``training.py:46-248``):

    python experiment_scripts/train_realestate10k.py --experiment_name vis --batch_size 12 --gpus 4 --synthetic 64

One process per GPU, Adam(lr, betas=(0.99, 0.999)), L1 image loss (+ ``--depth`` variance term),
gradient averaging across ranks (ONE flat all-reduce instead of the reference's per-parameter
loop), ``clip_grad_norm_(1.0)``, ``model_current.pth`` / ``model_final.pth`` checkpoints in the
reference's ``{'model','optimizer'}`` format.  The renderer's forward and backward run in the CUDA
library; the encoder and the optimizer are torch.
"""
import os
import shutil
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _common as C                                                   # noqa: E402
from cross_attention_renderer_b200.optim import FlatAdam             # noqa: E402
from cross_attention_renderer_b200.summaries import img_summaries    # noqa: E402


def multigpu_train(gpu, opt):
    dev = C.init_distributed(gpu, opt.gpus, opt.master_port)
    model = C.build_model(opt, dev)
    if opt.checkpoint_path is not None:
        C.load_checkpoint(model, opt.checkpoint_path, allow_encoder_mismatch=opt.allow_encoder_mismatch)
    if opt.gpus > 1:
        C.sync_model(model)
    # Adam(lr, betas=(0.99, 0.999)) of the reference (train_realestate10k.py:86) on one flat parameter buffer:
    # one all-reduce, one clip scalar and one fused kernel per step (cross_attention_renderer_b200/optim.py)
    optimizer = FlatAdam(model.parameters(), lr=opt.lr, betas=(0.99, 0.999))
    model.train()
    model.pixel_val_to_cpu = False

    root = os.path.join(opt.logging_root, opt.experiment_name)
    ckpt_dir = os.path.join(root, "checkpoints")
    if gpu == 0:
        if os.path.exists(root):
            shutil.rmtree(root)                                       # training.py:61-63 (overwrite=True)
        os.makedirs(ckpt_dir)
    writer = None
    if gpu == 0:
        try:
            from torch.utils.tensorboard import SummaryWriter
            writer = SummaryWriter(os.path.join(root, "summaries"))    # training.py:77
        except Exception as exc:                                       # noqa: BLE001
            print(f"tensorboard unavailable ({exc!r}): no summaries", flush=True)
    if not opt.synthetic:
        raise RuntimeError("RealEstate10k frames are not available in this environment: pass --synthetic N "
                           "(the reference loader is dataset/realestate10k_dataio.py)")
    rays = 192                                                        # query_sparsity (train_realestate10k.py:78)
    steps_per_epoch = max(1, opt.synthetic // opt.batch_size)
    total_steps, t0 = 0, time.time()
    for epoch in range(opt.num_epochs):
        for step in range(steps_per_epoch):
            # every rank draws its own batch (the reference builds an independent shuffled loader per rank)
            seed = (epoch * steps_per_epoch + step) * opt.gpus + gpu
            model_input, gt = C.synthetic_scene_batch(opt.batch_size, opt.sidelength, seed, rays=rays, device=dev)
            model_output = model(model_input)                         # z=None -> get_z inside (training.py:92)
            train_loss = C.image_loss(model_output, gt)
            if opt.depth and model_output["depth_ray"].shape[1] % 1024 == 0:
                train_loss = train_loss + C.depth_variance_loss(model_output, opt.l2_coeff)
            if not total_steps % opt.steps_til_summary and gpu == 0:
                C.save_checkpoint(model, optimizer, os.path.join(ckpt_dir, "model_current.pth"))
                if writer is not None:
                    # training.py:105-120,142-231: loss scalars + a full-image validation render with image summaries
                    writer.add_scalar("img_loss", float(train_loss.detach()), total_steps)
                    writer.add_scalar("total_at_entropy", C.attention_entropy(model_output["at_wt"]), total_steps)
                    model.eval()
                    with torch.no_grad():
                        val_in, val_gt = C.synthetic_scene_batch(1, opt.sidelength, 900_000 + total_steps, device=dev)
                        val_out = model(val_in)
                        val_out["pixel_val"] = val_out["pixel_val"].cpu()
                        img_summaries(model, val_in, val_gt, None, val_out, writer, total_steps, prefix="val_",
                                      img_shape=(opt.sidelength, opt.sidelength), n_view=opt.views)
                        writer.add_scalar("val_loss", float(C.image_loss(val_out, val_gt)), total_steps)
                    model.train()
                print(f"step {total_steps}: loss {float(train_loss.detach()):.5f} "
                      f"at_entropy {C.attention_entropy(model_output['at_wt']):.4f} "
                      f"({time.time() - t0:.1f} s)", flush=True)
            optimizer.zero_grad()
            train_loss.backward()
            # average_gradients + clip_grad_norm_(1.0) + Adam.step (training.py:21-28,127-136)
            optimizer.step(max_grad_norm=1.0)
            total_steps += 1
            if opt.iters_til_ckpt and not total_steps % opt.iters_til_ckpt and gpu == 0:
                C.save_checkpoint(model, optimizer,
                                  os.path.join(ckpt_dir, "model_epoch_%04d_iter_%06d.pth" % (epoch, total_steps)))
            if opt.max_steps and total_steps >= opt.max_steps:
                break
        if opt.max_steps and total_steps >= opt.max_steps:
            break
    if gpu == 0:
        if writer is not None:
            writer.close()
        C.save_checkpoint(model, optimizer, os.path.join(ckpt_dir, "model_final.pth"))
        print(f"done: {total_steps} steps, final loss {float(train_loss.detach()):.5f}", flush=True)
    if opt.gpus > 1:
        torch.distributed.destroy_process_group()
    return float(train_loss.detach())


def main(argv=None):
    opt = C.base_parser(__doc__, train=True).parse_args(argv)
    C.spawn(multigpu_train, opt)


if __name__ == "__main__":
    main()
