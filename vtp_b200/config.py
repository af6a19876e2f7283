"""VTPConfig — same fields, defaults and `model_type` as the reference's
vtp/models/vtp_hf/configuration_vtp.py:65-166 so that a reference `config.json` loads unchanged."""
from __future__ import annotations

from typing import Optional

try:  # transformers is only used as the (de)serialisation container, exactly like the reference
    from transformers import PretrainedConfig as _Base
except Exception:  # pragma: no cover - transformers is present in the target image
    class _Base:  # minimal stand-in
        def __init__(self, **kwargs):
            for k, v in kwargs.items():
                setattr(self, k, v)

        def to_dict(self):
            return dict(self.__dict__)


class VTPConfig(_Base):
    model_type = "vtp"

    def __init__(
        self,
        # General (configuration_vtp.py:69-72)
        image_size: int = 256,
        train_clip: bool = True,
        train_reconstruction: bool = True,
        # Vision encoder (:74-85)
        vision_patch_size: int = 16,
        vision_embed_dim: int = 768,
        vision_depth: int = 12,
        vision_num_heads: int = 12,
        vision_mlp_ratio: float = 4.0,
        vision_ffn_layer: str = "swiglu",
        vision_norm_layer: str = "rmsnorm",
        vision_init_values: Optional[float] = None,
        vision_use_qk_norm: bool = False,
        vision_feature_bottleneck: int = 64,
        vision_bottleneck_ae_only: bool = True,
        vision_clip_feat: str = "cls",
        # Text encoder (:87-101)
        text_context_length: int = 77,
        text_vocab_size: int = 49408,
        text_embed_dim: int = 768,
        text_num_heads: int = 12,
        text_depth: int = 12,
        text_mlp_ratio: float = 4.0,
        text_ls_init_value: Optional[float] = None,
        text_embed_cls: bool = False,
        text_pad_id: int = 0,
        text_no_causal_mask: bool = False,
        text_pool_type: str = "argmax",
        text_proj_type: str = "linear",
        text_proj_bias: bool = False,
        text_output_tokens: bool = False,
        text_quick_gelu: bool = False,
        # Pixel decoder (:103-110)
        decoder_embed_dim: int = 768,
        decoder_num_heads: int = 12,
        decoder_depth: int = 12,
        decoder_ffn_layer: str = "swiglu",
        decoder_norm_layer: str = "layernorm",
        decoder_init_values: Optional[float] = None,
        decoder_use_qk_norm: bool = False,
        # Runtime (:112-114)
        init_logit_scale: Optional[float] = None,
        init_logit_bias: Optional[float] = None,
        nonscalar_logit_scale: bool = False,
        **kwargs,
    ):
        super().__init__(**kwargs)
        loc = dict(locals())
        for k in ("self", "kwargs", "__class__"):
            loc.pop(k, None)
        for k, v in loc.items():
            setattr(self, k, v)

    @classmethod
    def from_vtp_yaml(cls, yaml_path: str) -> "VTPConfig":
        """configuration_vtp.py:169-233 — build from a legacy VTP YAML (plain PyYAML instead of OmegaConf)."""
        import yaml

        with open(yaml_path) as f:
            cfg = yaml.safe_load(f)
        v, t, d = cfg["vtp_model"]["vision_encoder"], cfg["vtp_model"]["text_encoder"], cfg["vtp_model"]["pixel_decoder"]
        tr = cfg["training"]
        return cls(
            image_size=cfg["data"]["image_size"], train_clip=tr["train_clip"],
            train_reconstruction=tr["train_reconstruction"],
            vision_patch_size=v["patch_size"], vision_embed_dim=v["embed_dim"], vision_depth=v["depth"],
            vision_num_heads=v["num_heads"], vision_mlp_ratio=v["mlp_ratio"], vision_ffn_layer=v["ffn_layer"],
            vision_norm_layer=v["norm_type"], vision_init_values=v.get("init_values"),
            vision_use_qk_norm=v.get("use_qk_norm", False), vision_feature_bottleneck=v["vit_feature_bottleneck"],
            vision_bottleneck_ae_only=v["bottleneck_ae_only"], vision_clip_feat=v["clip_feat"],
            text_context_length=t["context_length"], text_vocab_size=t["vocab_size"], text_embed_dim=t["embed_dim"],
            text_num_heads=t["heads"], text_depth=t["layers"], text_mlp_ratio=t["mlp_ratio"],
            text_ls_init_value=t.get("ls_init_value"), text_embed_cls=t["embed_cls"], text_pad_id=t["pad_id"],
            text_no_causal_mask=t["no_causal_mask"], text_pool_type=t["pool_type"], text_proj_type=t["proj_type"],
            text_proj_bias=t["proj_bias"], text_output_tokens=t["output_tokens"], text_quick_gelu=t["quick_gelu"],
            decoder_embed_dim=d["embed_dim"], decoder_num_heads=d["num_heads"], decoder_depth=d["depth"],
            decoder_ffn_layer=d["ffn_layer"], decoder_norm_layer=d["norm_layer"],
            decoder_init_values=d.get("layerscale_init"), decoder_use_qk_norm=d.get("use_qk_norm", False),
            init_logit_scale=tr.get("init_logit_scale"), init_logit_bias=tr.get("init_logit_bias"),
            nonscalar_logit_scale=tr.get("nonscalar_logit_scale", False),
        )


# The S/B/L hyper-parameters are not in the reference repo (they live in each HF checkpoint's config.json);
# these are the SURVEY.md §8(d) ASSUMED bench configs built from vision_transformer.py:328-361 /
# pixel_decoder.py:166-214 factory triplets.
def preset(name: str, **over) -> VTPConfig:
    name = name.lower()
    if name in ("small", "s"):
        kw = dict(vision_embed_dim=384, vision_depth=12, vision_num_heads=6, text_embed_dim=384, text_num_heads=6,
                  text_depth=12, decoder_embed_dim=384, decoder_num_heads=6, decoder_depth=12)
    elif name in ("base", "b"):
        kw = dict()
    elif name in ("large", "l"):
        kw = dict(vision_embed_dim=1024, vision_depth=24, vision_num_heads=16, text_embed_dim=768, text_num_heads=12,
                  text_depth=12, decoder_embed_dim=1024, decoder_num_heads=16, decoder_depth=24)
    elif name == "tiny":  # test-only
        kw = dict(vision_embed_dim=128, vision_depth=2, vision_num_heads=2, text_embed_dim=128, text_num_heads=2,
                  text_depth=2, decoder_embed_dim=128, decoder_num_heads=2, decoder_depth=2, text_vocab_size=1000)
    else:
        raise ValueError(f"unknown preset {name}")
    kw.update(over)
    return VTPConfig(**kw)
