#!/usr/bin/env python
"""Trajectory-rendering driver with the reference's entry point and flags
(``experiment_scripts/render_realestate10k_traj.py``):

    python experiment_scripts/render_realestate10k_traj.py --experiment_name vis_realestate --gpus 1 \
        --views 2 --checkpoint_path model.pth --synthetic 1 --frames 16

Per scene: ``get_z`` once, then one ``forward`` per query pose of the trajectory (the reference
loops over 8192-ray chunks per frame, render_realestate10k_traj.py:96,118-130; the library chunks
internally).  The query camera moves between the two context cameras (the reference interpolates
the dataset's intermediate frames; with synthetic scenes the path is a straight line plus the small
circle of ``make_circle``, :76-82).  Frames are written as PNGs under ``vis/<scene>/`` (OpenCV; the
reference writes an mp4 through imageio, absent here) and the rays/s of the render loop is printed.
With ``--gpus N`` the rays of every frame are sharded across ranks and the tiles gathered on all.
"""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _common as C                                                   # noqa: E402


def make_circle(n, radius=0.1):
    angles = np.linspace(0, 4 * np.pi, n)
    return np.stack([np.cos(angles) * radius, np.sin(angles) * radius, np.zeros(n)], axis=-1)


def trajectory(ctx_c2w, n):
    """(n,4,4) query poses: translation from context camera 0 to context camera 1 plus a small circle,
    orientation of the identity (the synthetic context pair converges on the -z axis)."""
    t0, t1 = ctx_c2w[0, :3, 3], ctx_c2w[1, :3, 3]
    w = torch.linspace(0.1, 0.9, n, device=ctx_c2w.device)[:, None]
    circ = torch.as_tensor(make_circle(n, 0.03), dtype=torch.float32, device=ctx_c2w.device)
    poses = torch.eye(4, device=ctx_c2w.device)[None].repeat(n, 1, 1)
    poses[:, :3, 3] = t0[None] * (1 - w) + t1[None] * w + circ
    return poses


def render(gpu, opt):
    dev = C.init_distributed(gpu, opt.gpus, opt.master_port)
    model = C.build_model(opt, dev)
    if opt.checkpoint_path is not None:
        C.load_checkpoint(model, opt.checkpoint_path, allow_encoder_mismatch=opt.allow_encoder_mismatch)
    model = model.eval()
    model.pixel_val_to_cpu = False
    if not opt.synthetic:
        raise RuntimeError("RealEstate10k frames are not available in this environment: pass --synthetic N")
    H = opt.sidelength
    out_root = os.path.join(opt.logging_root, "vis")
    try:
        import cv2
    except Exception:                                                 # noqa: BLE001
        cv2 = None
    results = []
    with torch.no_grad():
        for scene in range(opt.synthetic):
            model_input, _ = C.synthetic_scene_batch(1, H, 20_000 + scene, device=dev)
            z = model.get_z(model_input)                              # render_realestate10k_traj.py:97
            poses = trajectory(model_input["context"]["cam2world"][0], opt.frames)
            frames, t_frame = [], []
            torch.cuda.synchronize()
            t0 = time.time()
            for i in range(opt.frames):
                model_input["query"]["cam2world"] = poses[None, i:i + 1]
                if opt.gpus > 1:
                    out = C.sharding.render_sharded(model, model_input, z)
                else:
                    out = model(model_input, z=z)
                frames.append(out["rgb"].view(H, H, 3))
                if opt.frame_times:
                    torch.cuda.synchronize()
                    t_frame.append(time.time())
            torch.cuda.synchronize()
            dt = time.time() - t0
            rps = opt.frames * H * H / dt
            results.append(rps)
            if gpu == 0:
                print(f"scene {scene}: {opt.frames} frames in {dt:.3f} s = {rps:,.0f} rays/s", flush=True)
                if opt.frame_times:
                    ms = [round(1e3 * (b_ - a_), 1) for a_, b_ in zip([t0] + t_frame[:-1], t_frame)]
                    print("  ms per frame:", ms, flush=True)
                scene_dir = os.path.join(out_root, f"scene_{scene:04d}")
                os.makedirs(scene_dir, exist_ok=True)
                for i, fr in enumerate(frames):
                    img = ((fr.clamp(-1, 1) + 1) / 2 * 255).to(torch.uint8).cpu().numpy()
                    if cv2 is not None:
                        cv2.imwrite(os.path.join(scene_dir, f"{i:04d}.png"), img[..., ::-1])
                    else:
                        np.save(os.path.join(scene_dir, f"{i:04d}.npy"), img)
    if opt.gpus > 1:
        torch.distributed.destroy_process_group()
    return results


def main(argv=None):
    p = C.base_parser(__doc__)
    p.add_argument("--frames", type=int, default=16)
    p.add_argument("--frame_times", action="store_true", help="synchronise after every frame and print its latency")
    opt = p.parse_args(argv)
    C.spawn(render, opt)


if __name__ == "__main__":
    main()
