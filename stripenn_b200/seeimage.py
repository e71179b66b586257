"""``stripenn seeimage`` — mirrors reference src/stripenn/seeimage.py:32-97: the saturated heat
map (red = 1, green = blue = clip(255*(M - A)/M, 0)/255) of one genomic window for each maxpixel
quantile, written as ``<out>_<position>_<q>qt.png``.

The window image is produced on the GPU (`snn_heatmap`, csrc/band.cu) from the chromosome's
resident band / pixel table; the maxpixel values come from the same quantile kernels as
``compute`` (`-s` included).  The reference renders through matplotlib (axes, title, 640x480
canvas); matplotlib is not a dependency here, so the PNG holds the image itself, one pixel per
bin pair, quantised the way ``imshow`` quantises a float RGB array ((255*v) truncated) — the
float image is bit-identical to the array the reference hands to ``imshow``
(tests/golden/seeimage.npz)."""
from __future__ import annotations

import struct
import sys
import zlib

import numpy as np

from . import getstripe as gs
from .contactmap import open_contact_map
from .engine import Device


def write_png(path: str, rgb: np.ndarray) -> None:
    """8-bit RGB PNG (filter 0 on every scan line, one IDAT)."""
    h, w, c = rgb.shape
    if c != 3 or rgb.dtype != np.uint8:
        raise ValueError("write_png wants an (h, w, 3) uint8 array")
    raw = np.empty((h, 1 + 3 * w), dtype=np.uint8)
    raw[:, 0] = 0
    raw[:, 1:] = rgb.reshape(h, 3 * w)

    def chunk(tag: bytes, data: bytes) -> bytes:
        return struct.pack(">I", len(data)) + tag + data + struct.pack(">I", zlib.crc32(tag + data) & 0xffffffff)

    with open(path, "wb") as f:
        f.write(b"\x89PNG\r\n\x1a\n")
        f.write(chunk(b"IHDR", struct.pack(">IIBBBBB", w, h, 8, 2, 0, 0, 0)))
        f.write(chunk(b"IDAT", zlib.compress(raw.tobytes(), 6)))
        f.write(chunk(b"IEND", b""))


def parse_position(position: str, chromsize: int):
    """cooler region string -> (chrom, start, end) in bp, half-open; a bare name is the whole
    chromosome."""
    position = str(position)
    if ":" in position:
        chrom, rest = position.rsplit(":", 1)
        s, e = rest.split("-")
        return chrom, int(s.replace(",", "")), int(e.replace(",", ""))
    return position, 0, int(chromsize)


def window_images(cm, position, maxpixel, norm, slow=False, seed=123456789, device=None, device_id=0,
                  want_channel=True, verbose=True):
    """The images of seeimage on an in-memory contact map -> [(file-name suffix, float64 channel
    or None, uint8 RGB)] per maxpixel value (seeimage.py:36-37, 41-90)."""
    qs = list(map(float, maxpixel.split(',')))
    possible = cm.bin_columns()
    if norm == 'None':
        bal = False
    elif norm not in possible:
        if verbose:
            print('Possible normalization methods are:')
            print('None')
            for n in range(3, len(possible)):
                print(possible[n])
            print("Invalid normalization method. Normalization method is forced to None")
        bal = False
    else:
        bal = norm
    all_names = list(cm.chromnames)
    chrom = position.split(':')[0]
    if chrom not in all_names:
        sys.exit("Invalid chromosome name.")
    all_sizes = np.asarray(cm.chromsizes)
    keep = np.where(all_sizes > 500000)[0]
    all_names = [all_names[i] for i in keep]
    all_sizes = all_sizes[keep]
    if len(all_names) == 0:
        sys.exit("Exit: All chromosomes are shorter than 50kb.")
    csize = int(cm.chromsizes[cm.chromnames.index(chrom)])
    own = device is None
    dev = device if device is not None else Device(device_id)
    try:
        lib = gs.DeviceMatrix(cm, bal, dev)
        obj = gs.getStripe(lib, int(cm.binsize), 10, 8, 2.5, all_names, [chrom], all_sizes,
                           np.asarray([csize]), 2, 3, seed)
        if slow:
            if verbose:
                print("#####...Slowly estimating Maximum pixel values...#####")
            MP = obj.getQuantile_slow(cm, [chrom], qs)
        else:
            MP = obj.getQuantile_original(cm, [chrom], qs)
        _, start, end = parse_position(position, csize)
        i0, i1 = gs._bins_of(start, end, csize, int(cm.binsize))
        cid = lib.ensure(chrom)
        out = []
        for i, q in enumerate(qs):
            chan, rgb = dev.heatmap(cid, i0, i1, float(MP[chrom][i]), want_channel=want_channel)
            out.append(("_" + position + "_" + str(q) + "qt" + '.png', chan, rgb))
    finally:
        if own:
            dev.close()
    return out


def seeimage(cool, position, maxpixel, norm, out, slow, seed, device_id=0):
    """Drop-in for seeimage.seeimage (seeimage.py:32)."""
    cm = open_contact_map(cool)
    for suffix, _chan, rgb in window_images(cm, position, maxpixel, norm, slow, seed, device_id=device_id,
                                            want_channel=False):
        write_png(out + suffix, rgb)
    return 0
