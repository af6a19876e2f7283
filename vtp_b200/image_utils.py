"""Host-side image preparation used by the generation tokenizer (vtp/utils/image_utils.py:5-31 in the reference: the
ADM centre crop).  Data loading is outside the hot path; this exists so `VTP_Tokenizer.img_transform` is self-contained."""
from __future__ import annotations


def center_crop_arr(pil_image, image_size: int):
    """Square centre crop of side `image_size`: box-filter halvings while the short side is at least twice the target,
    one bicubic resize bringing the short side to the target (rounded per axis), then the centred window
    (offsets floor((side - target) / 2))."""
    from PIL import Image

    img = pil_image
    while min(img.size) >= 2 * image_size:
        img = img.resize((img.size[0] // 2, img.size[1] // 2), resample=Image.BOX)
    k = image_size / min(img.size)
    img = img.resize((round(img.size[0] * k), round(img.size[1] * k)), resample=Image.BICUBIC)
    left = (img.size[0] - image_size) // 2
    top = (img.size[1] - image_size) // 2
    return img.crop((left, top, left + image_size, top + image_size))
