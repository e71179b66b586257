"""``stripenn score`` driver — mirrors reference src/stripenn/score.py:7-61."""
from __future__ import annotations

import numpy as np
import pandas as pd

from . import getstripe as gs
from .compute import resolve_norm
from .contactmap import ContactMap, open_contact_map
from .engine import Device, default_band_width


def score_chromosomes(cm: ContactMap):
    """score.py:37-46."""
    names = list(cm.chromnames)
    sizes = np.asarray(cm.chromsizes)
    ok = [i for i in range(len(names)) if names[i] != "Y" and "JH" not in names[i] and "RANDOM" not in names[i]]
    nm = [names[i] for i in ok]
    sz = sizes[ok]
    big = np.where(sz > 1000000)[0]
    return [nm[i] for i in big], sz[big]


def read_coordinates(coordinates):
    """score.py:11-22: header sniffing, first six columns renamed."""
    table = pd.read_csv(coordinates, header=None, sep='\t')
    el = table.iloc[0]
    if type(el[1]) == str or type(el[2]) == str:
        table = pd.read_csv(coordinates, header=0, sep='\t')
    table.columns = ['chr', 'pos1', 'pos2', 'chr2', 'pos3', 'pos4'] + table.columns[6:].tolist()
    return table


def score_table(cm: ContactMap, table: pd.DataFrame, norm='KR', numcores=1, seed=123456789, mask='0',
                device=None, device_id=0, verbose=False):
    """Stages of getScore on in-memory inputs; returns the table with the six added columns."""
    bal = resolve_norm(cm, norm, "score", verbose)
    names, sizes = score_chromosomes(cm)
    own = device is None
    dev = device if device is not None else Device(device_id)
    try:
        # user stripes can be longer than the search windows: widen the band to hold them
        r = int(cm.binsize)
        bs = int(50000 / r)
        span = int(((table['pos4'] - table['pos3']).abs().max() + (table['pos2'] - table['pos1']).abs().max()) // r)
        bw = max(default_band_width(r), min(span + 2 * bs + 16, 4096))
        lib = gs.DeviceMatrix(cm, bal, dev, band_width=bw)
        obj = gs.getStripe(lib, r, 10, 8, 2.5, names, names, sizes, sizes, numcores, 1, seed)
        EV = obj.mpmean()
        bg = obj.nulldist()
        table = table.copy()
        table['chr'] = table['chr'].astype(str)
        pval = obj.pvalue(*bg, table)
        table.insert(table.shape[1], 'pvalue_added', pval, True)
        MEAN, SUM = obj.getMean(table)
        s, MEANOE, TOTALOE = obj.scoringstripes(table, EV, mask)
        table.insert(table.shape[1], 'Stripiness_added', s, True)
        table.insert(table.shape[1], "O_Mean_added", MEAN, True)
        table.insert(table.shape[1], "O_Sum_added", SUM, True)
        table.insert(table.shape[1], 'O/E_Mean_added', MEANOE, True)
        table.insert(table.shape[1], 'O/E_Total_added', TOTALOE, True)
    finally:
        if own:
            dev.close()
    return table


def getScore(cool, coordinates, norm, numcores, seed, out, mask='0', device_id=0):
    """Drop-in for score.getScore (score.py:7)."""
    print('Run score function')
    table = read_coordinates(coordinates)
    cm = open_contact_map(cool)
    table = score_table(cm, table, norm, numcores, seed, mask, device_id=device_id, verbose=True)
    table.to_csv(out, sep="\t", header=True, index=False)
