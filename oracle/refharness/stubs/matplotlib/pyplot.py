"""Stub of matplotlib.pyplot for the reference's ``seeimage`` (src/stripenn/seeimage.py:92-95):
records what the reference hands to ``imshow`` / ``savefig`` instead of rendering it, so the
image ARRAY can be compared.  No arithmetic lives here.
TEST INFRASTRUCTURE ONLY (oracle/refharness) — never imported by the product."""

CALLS = []          # one dict per savefig: {"img": ndarray, "title": str, "path": str}
_cur = {}


class _Figure:
    def savefig(self, path, *a, **k):
        rec = dict(_cur)
        rec["path"] = str(path)
        CALLS.append(rec)


class _Axes:
    figure = _Figure()


def subplot(*_a, **_k):
    return _Axes()


def imshow(img, *_a, **_k):
    _cur["img"] = img.copy()


def title(t, *_a, **_k):
    _cur["title"] = str(t)
