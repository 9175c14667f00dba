"""Tensorboard summaries of the training loop (reference summaries.py:15-141), consumed through the same
``writer.add_scalar / add_image`` calls and tag names: ``ent``, ``predictions``, ``depth_images``,
``context_images``, ``query_images``, ``epipolar_line``, ``out_min/out_max``, ``trgt_min/trgt_max``.

matplotlib is absent here; the ``jet`` colour map used for the depth image (summaries.py:38-40) is restated from
its segment data (a 256-entry look-up table, like ``matplotlib.cm.get_cmap('jet')``)."""
import numpy as np
import torch
import torchvision

_JET = {
    "red": ((0.0, 0.0), (0.35, 0.0), (0.66, 1.0), (0.89, 1.0), (1.0, 0.5)),
    "green": ((0.0, 0.0), (0.125, 0.0), (0.375, 1.0), (0.64, 1.0), (0.91, 0.0), (1.0, 0.0)),
    "blue": ((0.0, 0.5), (0.11, 1.0), (0.34, 1.0), (0.65, 0.0), (1.0, 0.0)),
}


def jet(x):
    """x in [0, 1] (any shape) -> (..., 3) colours, 256-entry LUT with matplotlib's index rule."""
    grid = np.linspace(0.0, 1.0, 256)
    lut = np.stack([np.interp(grid, [p[0] for p in _JET[c]], [p[1] for p in _JET[c]]) for c in ("red", "green", "blue")], -1)
    xa = np.asarray(x, dtype=np.float64)
    idx = np.where(xa >= 1.0, 255, np.clip((xa * 256).astype(np.int64), 0, 255))
    return lut[idx]


def _flatten_first_two(t):
    return t.reshape(-1, *t.shape[2:])


def img_summaries(model, model_input, ground_truth, loss_summaries, model_output, writer, iter, prefix="",
                  img_shape=(98, 144), n_view=1):
    predictions = model_output["rgb"]
    predictions = predictions.view(*predictions.size()[:-2], img_shape[0], img_shape[1], 3)
    predictions = _flatten_first_two(predictions).permute(0, 3, 1, 2)
    predictions = torch.clamp(predictions, -1, 1)
    if "at_wt" in model_output:
        at_wt = model_output["at_wt"]
        ent = -(at_wt * torch.log(at_wt + 1e-5)).sum(dim=-1).mean()
        writer.add_scalar(prefix + "ent", ent, iter)
    grid = lambda t, scale_each=False: torchvision.utils.make_grid(t, scale_each=scale_each, normalize=True).cpu().numpy()
    writer.add_image(prefix + "predictions", grid(predictions), iter)
    depth_img = model_output["depth_ray"].view(-1, img_shape[0], img_shape[1]).detach().cpu().numpy() / 10.0
    depth_img = torch.Tensor(jet(depth_img).transpose((0, 3, 1, 2)))
    writer.add_image(prefix + "depth_images", grid(depth_img, scale_each=True), iter)
    context_images = _flatten_first_two(model_input["context"]["rgb"]).permute(0, 3, 1, 2)
    writer.add_image(prefix + "context_images", grid(context_images), iter)
    query_images = model_input["query"]["rgb"]
    query_images = query_images.view(*query_images.size()[:-2], img_shape[0], img_shape[1], 3)
    query_images = _flatten_first_two(query_images).permute(0, 3, 1, 2)
    writer.add_image(prefix + "query_images", grid(query_images), iter)
    epi_summary(model_output, query_images, context_images, writer, iter, prefix=prefix, n_view=n_view)
    writer.add_scalar(prefix + "out_min", predictions.min(), iter)
    writer.add_scalar(prefix + "out_max", predictions.max(), iter)
    writer.add_scalar(prefix + "trgt_min", query_images.min(), iter)
    writer.add_scalar(prefix + "trgt_max", query_images.max(), iter)


def epi_summary(model_output, trgt_imgs_tile, ctxt_imgs_tile, writer, iter, prefix="", n_view=1):
    """One target pixel per scene, its epipolar samples in every context image and the sample with the largest
    round-1 attention weight (summaries.py:72-141).  The reference hard-codes ray 2065; clipped to the ray count."""
    pixel_val = model_output["pixel_val"].cpu().numpy()
    at_wt_max = model_output["at_wt_max"]
    uv = model_output["uv"]
    trgt_imgs_tile = trgt_imgs_tile.clone()
    ctxt_imgs_tile = ctxt_imgs_tile.clone()
    B, _, H, W = trgt_imgs_tile.size()
    s = pixel_val.shape
    pixel_val = pixel_val.reshape((s[0] // n_view, n_view, *s[1:]))
    s = at_wt_max.shape
    at_wt_max = at_wt_max.reshape((s[0] // n_view, n_view, *s[1:]))
    pix_size = H // 64 + 1
    counter = 0

    def box(x, y):
        return (max(x - pix_size, 0), min(x + pix_size, W - 1), max(y - pix_size, 0), min(y + pix_size, H - 1))
    for i in range(B):
        six = min(2065, uv.shape[2] - 1)
        coord = uv[i, 0, six]
        xmin, xmax, ymin, ymax = box(int(coord[0]), int(coord[1]))
        trgt_imgs_tile[i, :, ymin:ymax, xmin:xmax] = -1.0
        for k in range(n_view):
            for j in range(pixel_val.shape[3]):
                val = np.clip((pixel_val[i, k, six, j] + 1) / 2, 0, 1)
                xmin, xmax, ymin, ymax = box(int(val[0] * (W - 1)), int(val[1] * (H - 1)))
                ctxt_imgs_tile[counter, :, ymin:ymax, xmin:xmax] = 0.0
            max_idx = int(at_wt_max[i, k, six].item())
            val = np.clip((pixel_val[i, k, six, max_idx] + 1) / 2, 0, 1)
            xmin, xmax, ymin, ymax = box(int(val[0] * (W - 1)), int(val[1] * (H - 1)))
            ctxt_imgs_tile[counter, :, ymin:ymax, xmin:xmax] = -1.0
            counter += 1
    s = ctxt_imgs_tile.size()
    ctxt_imgs_tile = ctxt_imgs_tile.view(-1, n_view, *s[1:]).permute(1, 0, 2, 3, 4).reshape(*s)
    panel = torch.cat((trgt_imgs_tile, ctxt_imgs_tile), dim=0)
    writer.add_image(prefix + "epipolar_line",
                     torchvision.utils.make_grid(panel, scale_each=False, normalize=True).cpu().detach().numpy(), iter)
