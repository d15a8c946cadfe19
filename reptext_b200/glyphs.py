"""Host-side glyph conditioning of the RepText pipelines (stays Python / PIL / OpenCV, as in the reference:
``RepText/infer.py:16-21`` Canny, ``:64-104`` the per-line loop; ``RepText/infer_inpaint.py`` builds the same lists).

For every text line the reference renders the glyphs on a black canvas, takes the text bounding box as the position
image, the bounding box grown by 5 px as the regional mask, and an inverted Canny edge map of the glyph image as the
ControlNet image; the sum of all glyph images is the optional ``control_glyph`` init image.  This module restates that
loop as functions so that the drop-in pipelines can be driven exactly like ``infer.py`` drives the reference.

Arabic (and other right-to-left, contextually shaped) text must be reshaped and reordered BEFORE PIL draws it unless PIL
was built with libraqm: ``shape_text`` leaves the string alone when raqm is available, uses ``arabic_reshaper`` +
``python-bidi`` when they are installed, and otherwise the built-in shaper of ``reptext_b200/arabic.py`` (presentation
forms from the Unicode Character Database, lam-alef ligatures, UBA reordering for mixed Arabic / digits / Latin lines).
"""
from __future__ import annotations

import re
from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

try:  # host-only dependencies, as in the reference
    import cv2
except Exception:  # pragma: no cover
    cv2 = None
from PIL import Image, ImageDraw, ImageFont, features

_RTL = re.compile("[\u0590-\u08FF\uFB1D-\uFDFF\uFE70-\uFEFF]")
_CJK = re.compile("[\u4e00-\u9fff]")


def contains_chinese(text: str) -> bool:
    """``infer.py:11-14``: Chinese glyph lines are not appended to the prompt."""
    return bool(_CJK.search(text))


def contains_rtl(text: str) -> bool:
    return bool(_RTL.search(text))


def shape_text(text: str) -> str:
    """Visual-order, contextually shaped string for PIL's basic layout engine (identity for left-to-right text)."""
    if not contains_rtl(text):
        return text
    if features.check("raqm"):
        return text  # PIL shapes and reorders itself
    try:
        import arabic_reshaper
        from bidi.algorithm import get_display
        return get_display(arabic_reshaper.reshape(text))
    except Exception:
        from . import arabic
        return arabic.shape(text)


def canny(img_bgr: np.ndarray, low_threshold: int = 50, high_threshold: int = 100) -> np.ndarray:
    """``infer.py:16-21``: Canny edges, replicated to 3 channels and inverted (black edges on white)."""
    if cv2 is None:
        raise RuntimeError("OpenCV is required for the Canny condition (as in the reference)")
    e = cv2.Canny(img_bgr, low_threshold, high_threshold)[:, :, None]
    return 255 - np.concatenate([e, e, e], axis=2)


def load_font(font_path: Optional[str], font_size: int):
    """``infer.py:39-41``; without a font file PIL's bundled default face is used (no system fonts on the GPU boxes)."""
    if font_path:
        return ImageFont.truetype(font_path, font_size)
    return ImageFont.load_default(font_size)


@dataclass
class GlyphConditions:
    """The four arguments ``infer.py:117-130`` passes to the pipeline."""
    control_image: List[Image.Image] = field(default_factory=list)      # inverted Canny of each line
    control_position: List[Image.Image] = field(default_factory=list)   # text bbox, 0 / 255
    control_mask: List[Image.Image] = field(default_factory=list)       # bbox grown by `mask_margin`, 0 / 255
    control_glyph: Optional[Image.Image] = None                         # sum of the glyph images
    bboxes: List[Tuple[int, int, int, int]] = field(default_factory=list)


def resize_img(input_image: Image.Image, max_side: int = 1280, min_side: int = 1024, size=None,
               pad_to_max_side: bool = False, mode=Image.BILINEAR, base_pixel_number: int = 64) -> Image.Image:
    """``infer_inpaint.py:25-46``: the source photograph of the inpaint script brought to the model's sizes - the short side
    to ``min_side``, then the long side to ``max_side`` (two resizes, as upstream), both sides rounded DOWN to multiples of
    ``base_pixel_number``; ``size`` overrides; ``pad_to_max_side`` centres the result on a white square."""
    w, h = input_image.size
    if size is not None:
        w_new, h_new = size
    else:
        ratio = min_side / min(h, w)
        w, h = round(ratio * w), round(ratio * h)
        ratio = max_side / max(h, w)
        input_image = input_image.resize([round(ratio * w), round(ratio * h)], mode)
        w_new = (round(ratio * w) // base_pixel_number) * base_pixel_number
        h_new = (round(ratio * h) // base_pixel_number) * base_pixel_number
    input_image = input_image.resize([w_new, h_new], mode)
    if pad_to_max_side:
        canvas = np.ones([max_side, max_side, 3], dtype=np.uint8) * 255
        ox, oy = (max_side - w_new) // 2, (max_side - h_new) // 2
        canvas[oy:oy + h_new, ox:ox + w_new] = np.array(input_image)
        input_image = Image.fromarray(canvas)
    return input_image


def build_conditions(text_list: Sequence[str], text_position_list: Sequence[Tuple[int, int]],
                     text_color_list: Sequence[Tuple[int, int, int]], width: int, height: int, font,
                     mask_margin: int = 5, position_margin: int = 0) -> GlyphConditions:
    """The per-line loop of ``infer.py:64-104`` (one ControlNet pass per line in the pipeline).  ``position_margin``: the
    inpaint script grows the POSITION box by 5 pixels as well (``infer_inpaint.py:99``); ``infer.py`` does not."""
    if not (len(text_list) == len(text_position_list) == len(text_color_list)):
        raise ValueError("text_list, text_position_list and text_color_list must have the same length")
    out = GlyphConditions()
    glyph_all = np.zeros([height, width, 3], dtype=np.uint8)
    for text, pos, color in zip(text_list, text_position_list, text_color_list):
        shown = shape_text(text)
        glyph = Image.new("RGB", (width, height), (0, 0, 0))
        draw = ImageDraw.Draw(glyph)
        draw.text(pos, shown, font=font, fill=tuple(color))
        x0, y0, x1, y1 = draw.textbbox(pos, shown, font=font)
        x0, y0, x1, y1 = max(x0, 0), max(y0, 0), min(x1, width), min(y1, height)
        out.bboxes.append((x0, y0, x1, y1))
        position = np.zeros([height, width], dtype=np.uint8)
        position[max(y0 - position_margin, 0):y1 + position_margin, max(x0 - position_margin, 0):x1 + position_margin] = 255
        out.control_position.append(Image.fromarray(position))
        mask = np.zeros([height, width], dtype=np.uint8)
        mask[max(y0 - mask_margin, 0):y1 + mask_margin, max(x0 - mask_margin, 0):x1 + mask_margin] = 255
        out.control_mask.append(Image.fromarray(mask))
        g = np.array(glyph)
        glyph_all += g  # uint8 wrap-around, like the reference's `control_glyph_all += control_glyph`
        edges = canny(g[:, :, ::-1].copy())
        out.control_image.append(Image.fromarray(edges[:, :, ::-1].copy()))
    out.control_glyph = Image.fromarray(glyph_all).convert("RGB")
    return out


def build_prompt(base_prompt: str, text_list: Sequence[str], suffix: str = "") -> str:
    """``infer.py:108-113``: non-Chinese lines are quoted into the prompt."""
    p = base_prompt
    for t in text_list:
        if not contains_chinese(t):
            p += f", '{t}'"
    return p + suffix
