"""Shared plumbing of the B200 drivers (mirrors of the reference's ``experiment_scripts/``).

What the reference scripts do around the hot path — flag parsing (configargparse flag sets,
train_realestate10k.py:22-57), model construction and checkpoint loading (``{'model','optimizer'}``
files, ``strict=False``; train_realestate10k.py:90-106, training.py:118-120), one process per GPU
(train_realestate10k.py:60-73,132-133), PSNR (eval_realestate10k.py:74-75,181-187) — restated on
argparse + torch only.  The RealEstate10k loaders (``dataset/realestate10k_dataio.py``) are used
when that package and its data are importable; otherwise ``--synthetic`` scenes (random smooth
images, wide-baseline cameras) stand in.  The image encoder is the reference's multi-view DPT-hybrid
(``cross_attention_renderer_b200/encoder.py``, ``--encoder dpt_hybrid``, the default) or the small declared
stand-in of ``standin_encoder.py`` (``--encoder standin``) for smoke runs.
"""
import argparse
import importlib
import math
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from cross_attention_renderer_b200 import sharding, synthetic                      # noqa: E402
from cross_attention_renderer_b200.models import CrossAttentionRenderer             # noqa: E402
from cross_attention_renderer_b200.standin_encoder import StandInEncoder            # noqa: E402


def base_parser(description, train=False):
    """The reference's flag names and defaults; flags that only select reference code that is
    not part of this repo are accepted and ignored (``--network``, ``--category`` … are dead in
    the reference too)."""
    p = argparse.ArgumentParser(description=description)
    p.add_argument("-c", "--config_filepath", required=False)
    p.add_argument("--logging_root", type=str, default=os.environ.get("CAR_LOGGING_ROOT", "./logs"))
    p.add_argument("--data_root", type=str, default=None)
    p.add_argument("--val_root", type=str, default=None)
    p.add_argument("--network", type=str, default="relu")
    p.add_argument("--category", type=str, default="donut")
    p.add_argument("--conditioning", type=str, default="hyper")
    p.add_argument("--experiment_name", type=str, required=True)
    p.add_argument("--num_context", type=int, default=0)
    p.add_argument("--batch_size", type=int, default=12 if train else 48)
    p.add_argument("--max_num_instances", type=int, default=None)
    p.add_argument("--num_trgt", type=int, default=1)
    p.add_argument("--views", type=int, default=2)
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--lr", type=float, default=5e-5 if train else 5e-4)
    p.add_argument("--l2_coeff", type=float, default=0.05)
    p.add_argument("--num_epochs", type=int, default=40001)
    p.add_argument("--lpips", action="store_true", default=False)
    p.add_argument("--depth", action="store_true", default=False)
    p.add_argument("--model", type=str, default="midas_vit")
    p.add_argument("--epochs_til_ckpt", type=int, default=10)
    p.add_argument("--steps_til_summary", type=int, default=500)
    p.add_argument("--iters_til_ckpt", type=int, default=10000)
    p.add_argument("--checkpoint_path", default=None)
    p.add_argument("--reconstruct", action="store_true", default=False)
    for flag in ("no_multiview", "no_sample", "no_latent_concat", "no_data_aug", "no_high_freq"):
        p.add_argument("--" + flag, action="store_true", default=False)
    # additions of this repo
    p.add_argument("--synthetic", type=int, default=0, help="number of synthetic scenes (no dataset needed)")
    p.add_argument("--sidelength", type=int, default=256)
    p.add_argument("--npoints", type=int, default=64)
    p.add_argument("--precision", type=str, default="fp32", choices=["fp32", "bf16", "fp32_simt"])
    p.add_argument("--encoder", type=str, default="dpt_hybrid", choices=["dpt_hybrid", "standin"],
                   help="dpt_hybrid: the reference's multi-view DPT-hybrid encoder (encoder.py, the reference's "
                        "state_dict keys); standin: a small convolutional stand-in for smoke runs")
    p.add_argument("--encoder_module", type=str, default=None,
                   help="'pkg.mod:factory' returning an encoder module (overrides --encoder)")
    p.add_argument("--allow_encoder_mismatch", action="store_true", default=False,
                   help="load a checkpoint whose encoder.* tensors do not match this model's encoder (renderer weights only)")
    p.add_argument("--max_steps", type=int, default=None, help="stop after this many iterations / scenes")
    p.add_argument("--master_port", type=int, default=1493 if train else 1492)
    return p


def build_model(opt, device):
    if opt.encoder_module:
        mod, fn = opt.encoder_module.split(":")
        encoder = getattr(importlib.import_module(mod), fn)()
    elif opt.encoder == "standin":
        encoder = StandInEncoder()
    else:
        encoder = "dpt_hybrid"
    model = CrossAttentionRenderer(no_multiview=opt.no_multiview, no_sample=opt.no_sample,
                                   no_latent_concat=opt.no_latent_concat, no_high_freq=opt.no_high_freq,
                                   model=opt.model, n_view=opt.views, npoints=opt.npoints,
                                   precision=opt.precision, encoder=encoder)
    return model.to(device)


def load_checkpoint(model, path, optimizer=None, allow_encoder_mismatch=False):
    """Reference file format: ``{'model': state_dict, 'optimizer': state_dict}``, loaded with
    ``strict=False``; the optimizer state is NOT restored (train_realestate10k.py:95-106)."""
    print(f"Loading weights from {path}...")
    sd = torch.load(path, map_location="cpu")
    missing, unexpected = model.load_state_dict(sd["model"], strict=False)
    # strict=False hides mismatches: say what was not loaded.  Renderer layers the n_view=2 forward never
    # touches (latent_avg_*, update_val_merge) are expected to be fine either way; encoder weights are not.
    enc_unexpected = [k for k in unexpected if k.startswith("encoder.")]
    enc_missing = [k for k in missing if k.startswith("encoder.")]
    if unexpected:
        print(f"  checkpoint keys NOT used by this model ({len(unexpected)}): {list(unexpected)[:6]}{' ...' if len(unexpected) > 6 else ''}")
    if missing:
        print(f"  model parameters NOT in the checkpoint ({len(missing)}): {list(missing)[:6]}{' ...' if len(missing) > 6 else ''}")
    if (enc_unexpected or enc_missing) and not allow_encoder_mismatch:
        raise RuntimeError(
            f"{len(enc_unexpected)} encoder.* tensors of the checkpoint do not fit this model's encoder and "
            f"{len(enc_missing)} encoder parameters stay at their initial values: the renderer would run on "
            "features the checkpoint was not trained with.  Use the encoder the checkpoint was trained with (--encoder "
            "dpt_hybrid for the reference's, or --encoder_module <pkg.mod:factory>), or --allow_encoder_mismatch to load "
            "the renderer weights only.")
    return missing, unexpected


def save_checkpoint(model, optimizer, path):
    os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
    torch.save({"model": model.state_dict(), "optimizer": optimizer.state_dict() if optimizer else {}}, path)


def init_distributed(rank, world, port):
    """One process per GPU (reference: ``init_method='tcp://localhost:149x'``)."""
    if world > 1:
        dist.init_process_group(backend="nccl", init_method=f"tcp://127.0.0.1:{port}", world_size=world, rank=rank)
    torch.cuda.set_device(rank)
    return torch.device("cuda", rank)


def sync_model(model):
    """Reference ``sync_model`` (train_realestate10k.py:60-62) with one flat broadcast."""
    params = list(model.parameters())
    flat = torch.cat([p.data.reshape(-1) for p in params])
    dist.broadcast(flat, 0)
    off = 0
    for p in params:
        n = p.numel()
        p.data.copy_(flat[off:off + n].view_as(p))
        off += n


def synthetic_scene_batch(b, H, seed, rays=None, device="cpu"):
    """``(model_input, gt)`` in the RealEstate10k loader's layout (realestate10k_dataio.py:456-466):
    context rgb (b,2,H,W,3) in [-1,1], cameras, query uv (b,1,R,2) and the query ground-truth
    colours (b,1,R,3).  Images are smooth random fields; the "ground truth" is one more of them."""
    inp = synthetic.make_inputs(b, H, H, seed=seed, rays=rays)
    g = torch.Generator().manual_seed(7000 + seed)

    def smooth(n):
        low = torch.rand(n, 3, 8, 8, generator=g) * 2 - 1
        return torch.nn.functional.interpolate(low, size=(H, H), mode="bicubic", align_corners=False).clamp(-1, 1)
    inp["context"]["rgb"] = smooth(b * 2).reshape(b, 2, 3, H, H).permute(0, 1, 3, 4, 2).contiguous()
    tgt = smooth(b).permute(0, 2, 3, 1).reshape(b, H * H, 3)
    uv = inp["query"]["uv"][:, 0].long()
    idx = uv[..., 1] * H + uv[..., 0]
    rgb = torch.gather(tgt, 1, idx[..., None].expand(-1, -1, 3))[:, None]
    inp["query"]["rgb"] = rgb
    gt = {"rgb": rgb}
    return synthetic.to_device(inp, device), synthetic.to_device(gt, device)


def realestate_loader(split, views, **kw):
    """The reference's dataset, if its package and data are present on this machine."""
    try:
        mod = importlib.import_module("dataset.realestate10k_dataio")
    except Exception as exc:                                   # noqa: BLE001
        raise RuntimeError("dataset.realestate10k_dataio is not importable here: run with --synthetic N "
                           f"({exc!r})")
    return mod, split, views, kw


def image_loss(model_out, gt):
    """loss_functions.py:74-80: L1 on rgb with NaNs zeroed."""
    gt_rgb = torch.nan_to_num(gt["rgb"], nan=0.0)
    rgb = torch.nan_to_num(model_out["rgb"], nan=0.0)
    return torch.abs(gt_rgb - rgb).mean()


def depth_variance_loss(model_out, l2_weight):
    """loss_functions.py:120-129 (``--depth``): variance of the expected depth inside each 32x32
    ray patch (needs the ray count to be a multiple of 1024; mask of ones)."""
    depth_ray = model_out["depth_ray"][..., 0].reshape(-1, 1, 32, 32)
    mean = depth_ray.mean(dim=-1).mean(dim=-1)[:, None, None]
    return (l2_weight * torch.pow(depth_ray - mean, 2).mean(dim=-1).mean(dim=-1).mean(dim=-1)).mean()


def attention_entropy(at_wt):
    """training.py:110-114 (tensorboard scalar ``total_at_entropy``)."""
    ent = -(at_wt * torch.log(at_wt + 1e-5)).sum(dim=-1)
    return float(torch.nan_to_num(ent, nan=0.0).mean())


def psnr_masked(rgb, target, valid_mask):
    """eval_realestate10k.py:177-187: colours to [0,1], invalid rays grey on both sides, then
    ``-10 log10(mse)``."""
    rgb = ((rgb + 1) * 0.5) * valid_mask + 0.5 * (1 - valid_mask)
    target = ((target + 1) * 0.5) * valid_mask + 0.5 * (1 - valid_mask)
    mse = torch.mean((rgb - target) ** 2)
    return float(mse), float(-10.0 * math.log10(max(float(mse), 1e-20)))


def spawn(fn, opt):
    if opt.gpus > 1:
        import torch.multiprocessing as mp
        mp.spawn(fn, nprocs=opt.gpus, args=(opt,))
    else:
        fn(0, opt)
