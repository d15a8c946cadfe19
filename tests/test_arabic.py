"""The built-in Arabic shaper of the host glyph path (reptext_b200/arabic.py): presentation forms, ligatures, visual order."""
import unicodedata

from reptext_b200 import arabic as A
from reptext_b200 import glyphs


def cps(s):
    return [ord(c) for c in s]


def test_forms_table_is_the_unicode_database():
    assert A.FORMS["ب"] == ["ﺏ", "ﺐ", "ﺑ", "ﺒ"]               # beh: isolated, final, initial, medial
    assert A.FORMS["ا"][:2] == ["ﺍ", "ﺎ"] and A.FORMS["ا"][2:] == [None, None]   # alef joins to the right only
    assert A.joining_type("ب") == "D" and A.joining_type("ا") == "R" and A.joining_type("ء") == "U"
    assert A.joining_type("َ") == "T" and A.joining_type("a") == "U"          # fatha is transparent
    for base, forms in A.FORMS.items():
        for tag, f in zip(("<isolated>", "<final>", "<initial>", "<medial>"), forms):
            if f is not None:
                assert unicodedata.decomposition(f) == f"{tag} {ord(base):04X}"


def test_words_are_shaped_and_reversed():
    # meem-initial reh-final | hah-initial beh-medial alef-final, drawn right to left
    assert cps(A.shape("مرحبا")) == [0xFE8E, 0xFE92, 0xFEA3, 0xFEAE, 0xFEE3]
    # seen-initial, lam + alef -> final ligature, meem isolated (alef does not join forward)
    assert cps(A.shape("سلام")) == [0xFEE1, 0xFEFC, 0xFEB3]
    assert cps(A.shape("لا")) == [0xFEFB]
    assert cps(A.reshape("بالعالم")) == [0xFE91, 0xFE8E, 0xFEDF, 0xFECC, 0xFE8E, 0xFEDF, 0xFEE2]
    # a vowel mark does not break the join and stays on its letter in the visual string
    s = A.shape("بَب")
    assert cps(s) == [0xFE90, 0xFE91, 0x064E]


def test_mixed_direction_lines():
    assert A.shape("RepText 2025") == "RepText 2025"
    v = A.shape("مرحبا 123 abc")                     # right-to-left paragraph: Arabic at the right, Latin run at the left
    assert v.startswith("abc 123 ") and cps(v[-5:]) == [0xFE8E, 0xFE92, 0xFEA3, 0xFEAE, 0xFEE3]
    v = A.shape("Hello مرحبا world")                 # left-to-right paragraph with an embedded Arabic word
    assert v.startswith("Hello ") and v.endswith(" world") and cps(v[6:11]) == [0xFE8E, 0xFE92, 0xFEA3, 0xFEAE, 0xFEE3]
    assert A.shape("(مرحبا)") == "(" + A.shape("مرحبا") + ")"      # brackets are mirrored, so they still enclose
    assert A.shape("12:30 مساء")[-5:] == "12:30"      # numbers keep their order and sit left of the word that follows them
    assert A.shape("a\nمرحبا").split("\n")[0] == "a"


def test_glyph_path_uses_the_shaper():
    out = glyphs.shape_text("مرحبا بالعالم")
    from PIL import features
    if not features.check("raqm"):
        assert out != "مرحبا بالعالم" and all(0xFE70 <= ord(c) <= 0xFEFF or c == " " for c in out)
    assert glyphs.shape_text("RepText") == "RepText"
