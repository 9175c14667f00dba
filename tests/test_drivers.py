"""experiment_scripts/ drivers: flag surface and checkpoint format on CPU; tiny end-to-end runs on the GPU."""
import os
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "experiment_scripts"))
sys.path.insert(0, ROOT)

import _common as C                                                    # noqa: E402

HAS_GPU = torch.cuda.is_available()


def test_reference_flag_surface():
    """Every flag the reference's train / eval / render scripts define parses here
    (train_realestate10k.py:22-57, eval_realestate10k.py:34-70)."""
    argv = ["--experiment_name", "vis", "--batch_size", "12", "--gpus", "4", "--views", "2", "--lr", "5e-5",
            "--l2_coeff", "0.05", "--num_epochs", "3", "--lpips", "--depth", "--model", "midas_vit",
            "--epochs_til_ckpt", "10", "--steps_til_summary", "500", "--iters_til_ckpt", "10000",
            "--checkpoint_path", "x.pth", "--no_multiview", "--no_data_aug", "--no_high_freq",
            "--logging_root", "/tmp/x", "--network", "relu", "--category", "donut", "--conditioning", "hyper",
            "--num_context", "0", "--num_trgt", "1", "--data_root", "/d", "--val_root", "/v", "--reconstruct"]
    opt = C.base_parser("t", train=True).parse_args(argv)
    assert opt.gpus == 4 and opt.views == 2 and opt.no_multiview and opt.depth and opt.lr == 5e-5
    assert C.base_parser("t").parse_args(["--experiment_name", "e"]).batch_size == 48


def test_branch_flags_reach_the_model():
    opt = C.base_parser("t").parse_args(["--experiment_name", "e", "--views", "3", "--encoder", "standin"])
    m = C.build_model(opt, "cpu")
    assert m.n_view == 3 and m.general
    opt = C.base_parser("t").parse_args(["--experiment_name", "e", "--no_sample", "--encoder", "standin"])
    m = C.build_model(opt, "cpu")
    assert m.no_sample and m.general
    opt = C.base_parser("t").parse_args(["--experiment_name", "e", "--views", "4", "--encoder", "standin"])
    with pytest.raises(NotImplementedError):
        C.build_model(opt, "cpu")


def test_checkpoint_format_roundtrip(tmp_path):
    """Reference file format {'model','optimizer'} with the reference's parameter names, strict=False load."""
    assert C.base_parser("t").parse_args(["--experiment_name", "e"]).encoder == "dpt_hybrid"      # the reference's encoder
    opt = C.base_parser("t").parse_args(["--experiment_name", "e", "--encoder", "standin"])
    m = C.build_model(opt, "cpu")
    optim = torch.optim.Adam(m.parameters(), lr=1e-4, betas=(0.99, 0.999))
    path = str(tmp_path / "checkpoints" / "model_current.pth")
    C.save_checkpoint(m, optim, path)
    ck = torch.load(path, map_location="cpu")
    assert set(ck) == {"model", "optimizer"}
    for k in ("query_encode_latent.weight", "latent_value.bias", "phi.blocks.2.fc_1.bias", "conv_map.weight",
              "update_val_merge.weight", "latent_avg_query.weight", "encode_latent.weight"):
        assert k in ck["model"], k
    m2 = C.build_model(opt, "cpu")
    # a reference checkpoint carries encoder.* keys of the DPT encoder: extra / missing keys must not raise
    ck["model"]["encoder.pretrained.model.cls_token"] = torch.zeros(1)
    torch.save(ck, path)
    # ... but silently running the trained renderer on a different encoder is refused unless asked for
    with pytest.raises(RuntimeError, match="encoder"):
        C.load_checkpoint(m2, path)
    missing, unexpected = C.load_checkpoint(m2, path, allow_encoder_mismatch=True)
    assert "encoder.pretrained.model.cls_token" in unexpected
    assert torch.equal(m2.phi.lin_out.weight, m.phi.lin_out.weight)
    # the DPT-hybrid encoder carries the reference's encoder.* keys: such a checkpoint loads without any override
    opt3 = C.base_parser("t").parse_args(["--experiment_name", "e", "--encoder", "dpt_hybrid"])
    m3 = C.build_model(opt3, "cpu")
    assert m3.state_dict()["encoder.pretrained.model.cls_token"].shape == (1, 1, 768)
    path3 = str(tmp_path / "checkpoints" / "dpt.pth")
    torch.save({"model": m3.state_dict(), "optimizer": {}}, path3)
    missing, unexpected = C.load_checkpoint(C.build_model(opt3, "cpu"), path3)
    assert not list(missing) and not list(unexpected)


def test_synthetic_batch_layout_and_psnr():
    inp, gt = C.synthetic_scene_batch(2, 32, seed=3, rays=50)
    assert inp["context"]["rgb"].shape == (2, 2, 32, 32, 3) and inp["query"]["uv"].shape == (2, 1, 50, 2)
    assert gt["rgb"].shape == (2, 1, 50, 3) and float(inp["context"]["rgb"].abs().max()) <= 1.0
    x = torch.zeros(4, 4, 3)
    mse, psnr = C.psnr_masked(x, x + 0.2, torch.ones(4, 4, 1))
    assert abs(mse - 0.01) < 1e-6 and abs(psnr - 20.0) < 1e-3           # (0.2/2)^2 = 0.01 -> 20 dB
    mse, _ = C.psnr_masked(x, x + 0.2, torch.zeros(4, 4, 1))            # invalid rays are grey on both sides
    assert mse == 0.0


@pytest.mark.gpu
@pytest.mark.skipif(not HAS_GPU, reason="needs a CUDA device")
@pytest.mark.parametrize("encoder", ["standin", "dpt_hybrid"])
def test_train_eval_render_end_to_end(tmp_path, encoder):
    import train_realestate10k as T
    import eval_realestate10k as E
    import render_realestate10k_traj as R
    common = ["--experiment_name", "t", "--logging_root", str(tmp_path), "--sidelength", "64", "--synthetic", "4",
              "--encoder", encoder]
    T.main(common + ["--batch_size", "2", "--max_steps", "3", "--steps_til_summary", "1", "--lr", "1e-4"])
    ck = os.path.join(str(tmp_path), "t", "checkpoints", "model_final.pth")
    assert os.path.exists(ck) and os.path.exists(os.path.join(os.path.dirname(ck), "model_current.pth"))
    sd = torch.load(ck, map_location="cpu")["model"]
    assert all(torch.isfinite(v).all() for v in sd.values())
    # the renderer weights moved (gradients reached them through the CUDA backward)
    opt = C.base_parser("t").parse_args(["--experiment_name", "e"])
    torch.manual_seed(0)
    E.main(common + ["--checkpoint_path", ck, "--max_steps", "2"])
    R.main(common[:4] + ["--sidelength", "64", "--synthetic", "1", "--frames", "2", "--checkpoint_path", ck,
                         "--encoder", encoder])
    assert os.path.exists(os.path.join(str(tmp_path), "vis", "scene_0000"))
