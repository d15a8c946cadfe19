"""Built-in Arabic shaping and right-to-left reordering for the host glyph path.

The reference draws its glyph images with ``ImageDraw.text`` (``RepText/infer.py:75``), which produces correct Arabic only
when PIL was built with libraqm.  Without raqm PIL lays code points out left to right, one glyph per code point, so the
string has to arrive (a) in VISUAL order and (b) with every letter already replaced by its contextual presentation form
(isolated / final / initial / medial; Unicode blocks Arabic Presentation Forms-A / -B, which practically every Arabic font
maps).  ``arabic_reshaper`` + ``python-bidi`` do that when installed; this module does it with the standard library only:

* the letter -> forms table is derived at import time from ``unicodedata`` (the compatibility decompositions
  ``<initial> 0628`` ...), so it is the Unicode Character Database's, not a hand-typed one;
* joining follows the Unicode cursive-joining rules at the level these forms need: dual-joining letters (have an initial /
  medial form), right-joining letters (final form only), transparent marks (category Mn), everything else non-joining;
  lam + alef (plain, madda, hamza above / below) become the mandatory ligatures;
* reordering is rules L2 / L4 of the Unicode Bidirectional Algorithm over levels resolved from the characters' bidi
  classes (strong R / AL, strong L, numbers, neutrals taking the direction of their neighbours): enough for sign-board
  lines that mix Arabic words, numbers and Latin words; explicit embedding / override / isolate controls are not handled.
"""
from __future__ import annotations

import unicodedata
from typing import Dict, List, Optional, Tuple

_FORM_TAGS = {"<isolated>": 0, "<final>": 1, "<initial>": 2, "<medial>": 3}


def _build_tables() -> Tuple[Dict[str, List[Optional[str]]], Dict[Tuple[str, str], List[Optional[str]]]]:
    single: Dict[str, List[Optional[str]]] = {}
    pairs: Dict[Tuple[str, str], List[Optional[str]]] = {}
    for lo, hi in ((0xFB50, 0xFDFF), (0xFE70, 0xFEFF)):
        for cp in range(lo, hi + 1):
            dec = unicodedata.decomposition(chr(cp)).split()
            if not dec or dec[0] not in _FORM_TAGS:
                continue
            form, base = _FORM_TAGS[dec[0]], [chr(int(x, 16)) for x in dec[1:]]
            if len(base) == 1 and unicodedata.category(base[0]) == "Lo":
                single.setdefault(base[0], [None] * 4)[form] = chr(cp)
            elif len(base) == 2 and base[0] == "ل" and base[1] in "آأإا":
                pairs.setdefault((base[0], base[1]), [None] * 4)[form] = chr(cp)
    return single, pairs


FORMS, LAM_ALEF = _build_tables()
TATWEEL = "ـ"


def joining_type(ch: str) -> str:
    """'D' dual-joining, 'R' right-joining (joins the PREVIOUS letter only), 'T' transparent, 'U' non-joining."""
    if ch == TATWEEL:
        return "D"
    f = FORMS.get(ch)
    if f is not None:
        if f[2] is not None or f[3] is not None:
            return "D"
        if f[1] is not None:
            return "R"
        return "U"
    if unicodedata.category(ch) in ("Mn", "Me", "Cf"):
        return "T"
    return "U"


def reshape(text: str) -> str:
    """Logical-order string with every Arabic letter replaced by its contextual presentation form."""
    # 1. lam-alef ligatures (marks between the two letters are kept after the ligature)
    units: List[Tuple[str, Optional[Tuple[str, str]]]] = []       # (text, ligature key or None)
    chars = list(text)
    i = 0
    while i < len(chars):
        ch = chars[i]
        if ch == "ل":
            j = i + 1
            marks = []
            while j < len(chars) and joining_type(chars[j]) == "T":
                marks.append(chars[j])
                j += 1
            if j < len(chars) and (ch, chars[j]) in LAM_ALEF:
                units.append((ch + chars[j], (ch, chars[j])))
                units.extend((m, None) for m in marks)
                i = j + 1
                continue
        units.append((ch, None))
        i += 1

    def jt(u) -> str:
        return "R" if u[1] is not None else joining_type(u[0])      # a lam-alef ligature joins to the right only

    out = []
    n = len(units)
    for k, u in enumerate(units):
        t = jt(u)
        if t in ("T", "U") and u[1] is None and u[0] not in FORMS:
            out.append(u[0])
            continue
        # previous / next non-transparent neighbours
        p = k - 1
        while p >= 0 and jt(units[p]) == "T":
            p -= 1
        q = k + 1
        while q < n and jt(units[q]) == "T":
            q += 1
        joins_prev = p >= 0 and jt(units[p]) == "D" and t in ("D", "R")
        joins_next = q < n and jt(units[q]) in ("D", "R") and t == "D"
        forms = LAM_ALEF[u[1]] if u[1] is not None else FORMS.get(u[0])
        if forms is None:                                           # tatweel
            out.append(u[0])
            continue
        want = 3 if (joins_prev and joins_next) else 1 if joins_prev else 2 if joins_next else 0
        for cand in (want, 1 if want == 3 else 0, 0):
            if forms[cand] is not None:
                out.append(forms[cand])
                break
        else:
            out.append(u[0])
    return "".join(out)


_MIRROR = {"(": ")", ")": "(", "[": "]", "]": "[", "{": "}", "}": "{", "<": ">", ">": "<", "«": "»", "»": "«"}


def _levels(text: str, base: int) -> List[int]:
    cls = [unicodedata.bidirectional(c) for c in text]
    kind: List[Optional[str]] = []
    for c in cls:
        if c in ("R", "AL"):
            kind.append("R")
        elif c == "L":
            kind.append("L")
        elif c in ("EN", "AN"):
            kind.append("N")
        else:
            kind.append(None)            # neutral / weak separators / marks: resolved from the neighbours below
    n = len(text)
    # marks take the type of the character they sit on
    for i in range(n):
        if cls[i] == "NSM" and i > 0:
            kind[i] = kind[i - 1]
    # numbers: European digits after a left-to-right letter continue that run (rule W7)
    last_strong = "R" if base else "L"
    strong_before = []
    for i in range(n):
        strong_before.append(last_strong)
        if kind[i] in ("R", "L"):
            last_strong = kind[i]
    base_dir = "R" if base else "L"
    res: List[str] = []
    for i in range(n):
        k = kind[i]
        if k == "N":
            k = "L" if (cls[i] == "EN" and strong_before[i] == "L") else "N"
        res.append(k)
    # neutrals: same direction on both sides -> that direction, else the paragraph's (rules N1 / N2); numbers count as R
    i = 0
    while i < n:
        if res[i] is not None:
            i += 1
            continue
        j = i
        while j < n and res[j] is None:
            j += 1
        left = base_dir if i == 0 else ("R" if res[i - 1] in ("R", "N") else "L")
        right = base_dir if j == n else ("R" if res[j] in ("R", "N") else "L")
        fill = left if left == right else base_dir
        # a separator between two numbers stays with the numbers (1,5 / 12:30)
        if i > 0 and j < n and res[i - 1] == "N" and res[j] == "N" and all(cls[t] in ("CS", "ES") for t in range(i, j)):
            fill = "N"
        for t in range(i, j):
            res[t] = fill
        i = j
    lv = []
    for k in res:
        if base:
            lv.append(1 if k == "R" else 2)
        else:
            lv.append(0 if k == "L" else 1 if k == "R" else 2)
    return lv


def visual_order(text: str, base_rtl: Optional[bool] = None) -> str:
    """Reorder one line from logical to visual order (UBA rules L2 + L4).  ``base_rtl=None``: the first strong character
    decides (rules P2 / P3)."""
    if base_rtl is None:
        base_rtl = False
        for c in text:
            b = unicodedata.bidirectional(c)
            if b in ("R", "AL"):
                base_rtl = True
                break
            if b == "L":
                break
    lv = _levels(text, 1 if base_rtl else 0)
    chars = [(_MIRROR.get(c, c) if l % 2 else c) for c, l in zip(text, lv)]
    for level in range(max(lv, default=0), 0, -1):
        i = 0
        while i < len(chars):
            if lv[i] >= level:
                j = i
                while j < len(chars) and lv[j] >= level:
                    j += 1
                chars[i:j] = chars[i:j][::-1]
                i = j
            else:
                i += 1
    # a combining mark must FOLLOW its base in the visual string too (PIL stacks it on the previous glyph)
    out: List[str] = []
    pending: List[str] = []
    for c in chars:
        if unicodedata.category(c) in ("Mn", "Me"):
            pending.append(c)
        else:
            out.append(c)
            out.extend(reversed(pending))
            pending = []
    out.extend(pending)
    return "".join(out)


def shape(text: str, base_rtl: Optional[bool] = None) -> str:
    """Presentation forms + visual order, line by line: what PIL's basic layout needs to draw Arabic correctly."""
    return "\n".join(visual_order(reshape(line), base_rtl) for line in text.split("\n"))
