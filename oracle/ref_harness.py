"""TEST INFRASTRUCTURE ONLY.  Imports the *real* reference (MiniMax-AI/VTP, /root/reference) in this container so that
the oracle restatement (oracle/vtp_oracle.py) and the committed golden vectors (tests/golden/) can be pinned against it.
/root/reference does not exist on the GPU box: nothing under tests -m gpu / bench.py / smoke() may import this module.

The reference imports `omegaconf` at module top (vtp/models/vtp.py:27) which is not installed; an in-memory stub
(attribute-dict DictConfig + OmegaConf.create) is injected before import.
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("VTP_REFERENCE_ROOT", "/root/reference")


class _AttrDict(dict):
    def __getattr__(self, k):
        try:
            v = self[k]
        except KeyError as e:
            raise AttributeError(k) from e
        return v

    def __setattr__(self, k, v):
        self[k] = v

    def get(self, k, default=None):
        return self[k] if k in self else default


def to_attrdict(obj):
    if isinstance(obj, dict):
        return _AttrDict({k: to_attrdict(v) for k, v in obj.items()})
    if isinstance(obj, (list, tuple)):
        return [to_attrdict(v) for v in obj]
    return obj


def _install_omegaconf_stub():
    if "omegaconf" in sys.modules:
        return
    m = types.ModuleType("omegaconf")
    m.DictConfig = _AttrDict

    class OmegaConf:
        @staticmethod
        def create(d):
            return to_attrdict(d)

        @staticmethod
        def load(path):
            import yaml

            with open(path) as f:
                return to_attrdict(yaml.safe_load(f))

        @staticmethod
        def merge(a, b):
            out = _AttrDict(a)
            out.update(b)
            return out

        @staticmethod
        def from_cli(lst):
            return _AttrDict()

    m.OmegaConf = OmegaConf
    sys.modules["omegaconf"] = m


def available() -> bool:
    return os.path.isdir(os.path.join(REF_ROOT, "vtp"))


def import_reference():
    """Returns the reference `vtp` package (imported from REF_ROOT)."""
    if not available():
        raise RuntimeError(f"reference not found at {REF_ROOT}")
    _install_omegaconf_stub()
    if REF_ROOT not in sys.path:
        sys.path.insert(0, REF_ROOT)
    import vtp  # noqa

    return vtp


def legacy_vtp_config(*, embed_dim, depth, heads, text_dim, text_heads, text_layers, dec_dim, dec_depth, dec_heads,
                      image_size=256, head_out_dim=4096, head_hidden=2048, head_bottleneck=256, vocab=49408,
                      drop=0.0):
    """DictConfig tree for the legacy 3-objective meta-arch (vtp/models/vtp.py:119-273). No example YAML ships with
    the reference; keys are the ones the constructor reads."""
    cfg = dict(
        data=dict(image_size=image_size),
        training=dict(train_clip=True, train_dinov2=True, train_reconstruction=True, cast_dtype=None,
                      init_logit_scale=None, init_logit_bias=None, nonscalar_logit_scale=False,
                      clip_output_dict=True, clip_drop_rate=drop, ssl_drop_rate=drop, rec_drop_rate=drop),
        vtp_model=dict(
            vision_encoder=dict(model_type="dinov3", patch_size=16, embed_dim=embed_dim, depth=depth,
                                num_heads=heads, mlp_ratio=4.0, ffn_layer="swiglu", norm_type="rmsnorm",
                                init_values=None, vit_feature_bottleneck=64, bottleneck_ae_only=True,
                                clip_feat="cls"),
            text_encoder=dict(context_length=77, vocab_size=vocab, embed_dim=text_dim, heads=text_heads,
                              layers=text_layers, mlp_ratio=4.0, ls_init_value=None, embed_cls=False,
                              no_causal_mask=False, pad_id=0, pool_type="argmax", proj_type="linear",
                              proj_bias=False, output_tokens=False, quick_gelu=False, norm_kwargs=None,
                              act_kwargs=None),
            pixel_decoder=dict(model_type="dinov3", embed_dim=dec_dim, depth=dec_depth, num_heads=dec_heads,
                               upscale_factor=16, norm_layer="layernorm", ffn_layer="swiglu"),
            dino_head=dict(out_dim=head_out_dim, nlayers=3, hidden_dim=head_hidden, bottleneck_dim=head_bottleneck),
        ),
    )
    return to_attrdict(cfg)
