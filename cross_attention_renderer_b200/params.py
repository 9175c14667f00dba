"""Checkpoint ABI of the per-ray renderer: parameter names and shapes.

The names/shapes are the ``state_dict`` keys the reference module creates in
``CrossAttentionRenderer.__init__`` (reference models.py:96-145) for
``model='midas_vit'`` and ``n_view=2`` (``latent_dim = 512+64 = 576`` before
the encode layers halve it to 288, models.py:94,102-104).  ``encoder.*`` keys
belong to the image encoder and are outside the hot path.

Only tensors listed in ``HOT_PATH_PARAMS`` are read by the n_view=2 forward
(models.py:281-344, 487-565, 611); the ``latent_avg_*`` / ``update_val_merge``
layers exist in the checkpoint but are never touched by that branch.
"""
from collections import OrderedDict

LATENT_IN = 576          # 256 (z1) + 256 (z2) + 64 (z3), models.py:94
LATENT = 288             # models.py:104
HIDDEN = 128             # models.py:114
LOCAL = 16               # local_coords channels, models.py:528
D_IN_PHI_PER_VIEW = 9    # plucker(6) + origin(3), models.py:144,597


def renderer_param_shapes(n_view=2, num_hidden_units_phi=128, no_latent_concat=False):
    """Ordered {name: shape} of every non-encoder parameter (models.py:96-145)."""
    if n_view not in (1, 2, 3):
        raise NotImplementedError("the reference defines n_view in {1, 2, 3}")
    # n_view > 1: the per-sample encoder halves the latent (models.py:100-104); n_view == 1 keeps the
    # 576 channels and merges the 6 point channels with update_val_merge (models.py:107-108)
    LATENT = LATENT_IN // 2 if (n_view > 1 and not no_latent_concat) else LATENT_IN
    # no_latent_concat: no per-sample encoder, raw 576-channel features, V / K read one view (models.py:106-107,121-124)
    kv_in = LATENT if no_latent_concat else LATENT * n_view
    h = HIDDEN
    hp = num_hidden_units_phi
    s = OrderedDict()

    def conv2d(name, cin, cout, k=1):
        s[name + ".weight"] = (cout, cin, k, k)
        s[name + ".bias"] = (cout,)

    conv2d("conv_map", 3, 64, 7)                                  # models.py:96
    if no_latent_concat:
        conv2d("feature_map", LATENT_IN, LATENT_IN // 2)          # :107 (defined, never evaluated)
    elif n_view > 1:
        conv2d("query_encode_latent", LATENT_IN + 3, LATENT_IN)   # :102
        conv2d("query_encode_latent_2", LATENT_IN, LATENT)        # :103
        conv2d("update_val_merge", LATENT * 2 + 6, LATENT)        # :105
    else:
        conv2d("update_val_merge", LATENT + 6, LATENT)            # :108
    conv2d("latent_value", kv_in, LATENT)                         # :117 / :122
    conv2d("key_map", kv_in, h)                                   # :118 / :123
    conv2d("key_map_2", h, h)                                     # :119
    conv2d("query_embed", LOCAL, h)                               # :126
    conv2d("query_embed_2", h, h)                                 # :127
    conv2d("latent_avg_query", 9 + 16, h)                         # :130
    conv2d("latent_avg_query_2", h, h)                            # :131
    conv2d("latent_avg_key", LATENT, h)                           # :133
    conv2d("latent_avg_key_2", h, h)                              # :134
    conv2d("query_repeat_embed", 16 + 128, h)                     # :136
    conv2d("query_repeat_embed_2", h, h)                          # :137
    conv2d("latent_avg_repeat_query", 9 + 16 + 128, h)            # :139
    conv2d("latent_avg_repeat_query_2", h, h)                     # :140
    s["encode_latent.weight"] = (128, LATENT, 1)                  # Conv1d, :142
    s["encode_latent.bias"] = (128,)
    # phi = ResnetFC(n_view*9, n_blocks=3, d_out=3, d_latent=LATENT*n_view,
    #                d_hidden=hp)  (models.py:144-145, resnet_block_fc.py:65-130)
    s["phi.lin_in.weight"] = (hp, n_view * D_IN_PHI_PER_VIEW)
    s["phi.lin_in.bias"] = (hp,)
    s["phi.lin_out.weight"] = (3, hp)
    s["phi.lin_out.bias"] = (3,)
    for i in range(3):
        s[f"phi.blocks.{i}.fc_0.weight"] = (hp, hp)
        s[f"phi.blocks.{i}.fc_0.bias"] = (hp,)
        s[f"phi.blocks.{i}.fc_1.weight"] = (hp, hp)
        s[f"phi.blocks.{i}.fc_1.bias"] = (hp,)
    for i in range(3):
        s[f"phi.lin_z.{i}.weight"] = (hp, LATENT * n_view)
        s[f"phi.lin_z.{i}.bias"] = (hp,)
    return s


# Layers the n_view=2 forward actually evaluates.
HOT_PATH_LAYERS = (
    "query_encode_latent", "query_encode_latent_2", "latent_value",
    "key_map", "key_map_2", "query_embed", "query_embed_2",
    "query_repeat_embed", "query_repeat_embed_2", "encode_latent",
    "phi.lin_in", "phi.lin_out",
    "phi.blocks.0.fc_0", "phi.blocks.0.fc_1",
    "phi.blocks.1.fc_0", "phi.blocks.1.fc_1",
    "phi.blocks.2.fc_0", "phi.blocks.2.fc_1",
    "phi.lin_z.0", "phi.lin_z.1", "phi.lin_z.2",
)
HOT_PATH_PARAMS = tuple(
    f"{l}.{p}" for l in HOT_PATH_LAYERS for p in ("weight", "bias"))
