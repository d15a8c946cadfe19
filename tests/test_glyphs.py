"""Host glyph path (reptext_b200/glyphs.py, the loop of RepText/infer.py:64-104): shapes, value ranges and the
relations between the four conditions, on the CPU; the infer.py-shaped driver end to end on the GPU."""
import numpy as np
import pytest

from reptext_b200 import glyphs


def test_conditions_of_two_lines():
    W, H = 256, 192
    font = glyphs.load_font(None, 32)
    c = glyphs.build_conditions(["RepText", "B200"], [(20, 30), (40, 100)], [(255, 255, 255), (255, 0, 0)], W, H, font)
    assert len(c.control_image) == len(c.control_position) == len(c.control_mask) == 2
    for img, pos, mask, (x0, y0, x1, y1) in zip(c.control_image, c.control_position, c.control_mask, c.bboxes):
        assert img.size == pos.size == mask.size == (W, H) and img.mode == "RGB"
        p, m, e = np.array(pos), np.array(mask), np.array(img)
        assert set(np.unique(p)) <= {0, 255} and set(np.unique(m)) <= {0, 255}
        assert p[y0:y1, x0:x1].min() == 255 and p.sum() == 255 * (y1 - y0) * (x1 - x0)      # exactly the text bbox
        assert (m >= p).all() and m.sum() > p.sum()                                           # bbox grown by 5 px
        assert e.max() == 255 and e.min() == 0                                                # inverted edges exist
        assert (e[m == 0] == 255).all()                                                       # no edges outside the mask
    g = np.array(c.control_glyph)
    assert g.shape == (H, W, 3) and g.max() == 255
    assert (g[np.array(c.control_mask[0]) + np.array(c.control_mask[1]) == 0] == 0).all()     # glyphs live inside masks


def test_prompt_and_script_detection():
    assert glyphs.build_prompt("a sign", ["Shakker Labs", "哩布哩布"]) == "a sign, 'Shakker Labs'"
    assert glyphs.contains_rtl("مرحبا") and not glyphs.contains_rtl("hello")
    assert glyphs.shape_text("hello") == "hello"
    with pytest.raises(ValueError):
        glyphs.build_conditions(["a"], [], [], 64, 64, glyphs.load_font(None, 16))


def test_rtl_text_renders_something():
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        c = glyphs.build_conditions(["مرحبا"], [(10, 10)], [(255, 255, 255)], 128, 64, glyphs.load_font(None, 24))
    assert np.array(c.control_glyph).max() > 0


@pytest.mark.gpu
def test_infer_shaped_driver_runs_end_to_end(tmp_path):
    import sys, os
    import torch
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import infer
    out = tmp_path / "result.png"
    img = infer.main(["--config", "small", "--text", "RepText", "--text", "B200", "--steps", "2", "--out", str(out)])
    assert img.size == (256, 256) and out.exists()
    lat = infer.main(["--config", "tiny", "--text", "RepText", "--steps", "2", "--output-type", "latent", "--no-glyph-init"])
    assert torch.isfinite(lat.float()).all() and tuple(lat.shape) == (1, 256, 64)


@pytest.mark.gpu
def test_infer_inpaint_shaped_driver_runs_end_to_end(tmp_path):
    """``examples/infer_inpaint.py`` (the flow of RepText/infer_inpaint.py: synthetic photograph, grown position box, true CFG,
    second ControlNet) - the command verified on a B200 at the end of round 2."""
    import sys, os
    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "examples"))
    import infer_inpaint
    out = tmp_path / "result_inpaint.png"
    img = infer_inpaint.main(["--config", "small", "--steps", "3", "--text", "RepText", "--text", "B200", "--out", str(out)])
    assert img.size == (256, 256) and out.exists()
