"""``stripenn compute`` driver — mirrors reference src/stripenn/stripenn.py:64-163 stage by stage
(norm resolution :80-92, chromosome filter :96, -k handling :104-116, stage order :120-152,
output layout :149-159) on top of the B200 :class:`getStripe`.
"""
from __future__ import annotations

import errno
import os
import shutil
import sys
import time
import warnings

import numpy as np
import pandas as pd

from . import getstripe as gs
from .contactmap import ContactMap, open_contact_map
from .engine import Device


def makeOutDir(outdir, assume_yes=False):
    """stripenn.py:12-42 (interactive overwrite prompt kept; ``assume_yes`` answers Y)."""
    if outdir[-1] != '/':
        outdir += '/'
    if os.path.exists(outdir):
        if assume_yes:
            userinput = 'Y'
        else:
            print('\n%s exists. Do you want to remove all files and save new results in this folder? [Y/n]' % (outdir))
            userinput = input()
        if userinput in ('Y', 'y'):
            print('All directories and files in %s will be deleted.' % (outdir))
            for filename in os.listdir(outdir):
                file_path = os.path.join(outdir, filename)
                try:
                    if os.path.isfile(file_path) or os.path.islink(file_path):
                        os.unlink(file_path)
                    elif os.path.isdir(file_path):
                        shutil.rmtree(file_path)
                except Exception as e:
                    print('Failed to delete %s with the reason: %s' % (file_path, e))
        elif userinput in ('n', 'N'):
            print('Input another output directory. Exit.')
            sys.exit()
        else:
            print('Type Y or n.\nExit.')
            sys.exit()
    else:
        try:
            os.makedirs(outdir)
        except OSError as e:
            if e.errno != errno.EEXIST:
                raise


def addlog(cool, out, norm, chrom, canny, minL, maxW, maxpixel, numcores, pvalue, mask, bfilter):
    """stripenn.py:45-61 — same keys, same order."""
    if out[-1] != '/':
        out += '/'
    with open(out + "stripenn.log", 'w') as f:
        f.write('cool: ' + str(cool) + '\n')
        f.write('out: ' + out + '\n')
        f.write('norm: ' + str(norm) + '\n')
        f.write('chrom: ' + str(chrom) + '\n')
        f.write('canny: ' + str(canny) + '\n')
        f.write('minL: ' + str(minL) + '\n')
        f.write('maxW: ' + str(maxW) + '\n')
        f.write('maxpixel: ' + str(maxpixel) + '\n')
        f.write('num_cores: ' + str(numcores) + '\n')
        f.write('pvalue: ' + str(pvalue) + '\n')
        f.write('mask: ' + str(mask) + '\n')
        f.write('blur filter: ' + str(bfilter) + '\n')


def resolve_norm(cm: ContactMap, norm, mode="compute", verbose=True):
    """stripenn.py:82-92 / score.py:27-35."""
    possible = cm.bin_columns()
    if norm == 'None':
        return False
    if mode == "compute" and norm == 'weight':
        return True
    if norm not in possible:
        if verbose:
            print('Possible normalization methods are:')
            print('None')
            for n in range(3, len(possible)):
                print(possible[n])
            print("Invalid normalization method. Normalization method is forced to None")
        return False
    return norm


def compute_chromosomes(cm: ContactMap, chroms):
    """stripenn.py:94-116 -> (all_chromnames, all_chromsizes, chromnames, chromsizes)."""
    names = list(cm.chromnames)
    sizes = np.asarray(cm.chromsizes)
    keep = [i for i in range(len(names))
            if 'JH5' not in names[i] and 'GL4' not in names[i] and 'RANDOM' not in names[i]
            and names[i] != 'M' and names[i] != "chrM" and names[i] != "Y" and names[i] != "chrY"]
    all_names = [names[i] for i in keep]
    all_sizes = sizes[keep]
    if len(all_names) == 0:
        sys.exit("Exit: All chromosomes are shorter than 50kb.")
    chromnames, chromsizes = all_names, all_sizes
    if chroms[0] != 'all':
        idx, warnflag = [], False
        for item in chroms:
            if item in all_names:
                idx.append(all_names.index(item))
            else:
                warnings.warn('\nThere is no chromosomes called ' + str(item) +
                              ' in the provided .cool file or it is shorter than 50kb.')
                warnflag = True
        if warnflag:
            warnings.warn('\nThe possible chromosomes are: ' + ', '.join(all_names))
        chromnames = chroms
        chromsizes = all_sizes[idx]
    return all_names, all_sizes, chromnames, chromsizes


def run_stages(cm: ContactMap, norm='KR', chrom='all', canny=2.0, minL=10, maxW=8,
               maxpixel='0.95,0.96,0.97,0.98,0.99', numcores=1, pvalue=0.1, mask='0', slow=False,
               bfilter=3, seed=123456789, device=None, device_id=0, gray_variant=0, verbose=False,
               return_stages=False):
    """Stages 1-5 of ``compute`` on an in-memory contact map: everything between "matrix
    available" and "result tables on the host" (the span the Mpixels/s metric is defined on).
    Returns (result_unfiltered, result_filtered) DataFrames [+ the stage products]."""
    say = print if verbose else (lambda *a, **k: None)
    qs = list(map(float, maxpixel.split(','))) if isinstance(maxpixel, str) else list(maxpixel)
    chroms = chrom.split(',') if isinstance(chrom, str) else list(chrom)
    bal = resolve_norm(cm, norm, "compute", verbose)
    all_names, all_sizes, chromnames, chromsizes = compute_chromosomes(cm, chroms)
    own = device is None
    dev = device if device is not None else Device(device_id)
    try:
        lib = gs.DeviceMatrix(cm, bal, dev)
        obj = gs.getStripe(lib, cm.binsize, minL, maxW, canny, all_names, chromnames, all_sizes,
                           chromsizes, numcores, bfilter, seed, gray_variant)
        say('1. Maximum pixel value calculation ...')
        if slow:
            MP = obj.getQuantile_slow(cm, chromnames, qs)
        else:
            MP = obj.getQuantile_original(cm, chromnames, qs)
        say('2. Expected value calculation ...')
        EV = obj.mpmean()
        say('3. Background distribution estimation ...')
        bg = obj.nulldist()
        say('4. Finding candidate stripes from each chromosome ...')
        per_q = obj.extract_all(MP, qs, *bg)
        result_table = pd.concat([pd.DataFrame(columns=gs.RESULT_COLUMNS + ['pvalue'])] + per_q) \
            if any(len(t) for t in per_q) else pd.DataFrame(columns=gs.RESULT_COLUMNS + ['pvalue'])
        result_table = obj.RemoveRedundant(result_table, 'pvalue')
        say('5. Stripiness calculation ...')
        s = obj.scoringstripes(result_table, EV, mask)[0]
        result_table = result_table.drop(columns=['total', 'num', 'start', 'end', 'x', 'y', 'h', 'w', 'medpixel'])
        result_table.insert(result_table.shape[1], 'Stripiness', s, True)
        res_filter = result_table[result_table['pvalue'] < pvalue]
        res_filter = res_filter.sort_values(by=['Stripiness'], ascending=False)
    finally:
        if own:
            dev.close()
    if return_stages:
        return result_table, res_filter, dict(MP=MP, EV=EV, bg=bg, per_q=per_q)
    return result_table, res_filter


def compute(cool, out, norm, chrom, canny, minL, maxW, maxpixel, numcores, pvalue, mask, slow, bfilter,
            seed, device_id=0, assume_yes=False):
    """Drop-in for stripenn.compute (stripenn.py:64): same arguments, same files written."""
    np.seterr(divide='ignore', invalid='ignore')
    t_start = time.time()
    if out[-1] != '/':
        out += '/'
    makeOutDir(out, assume_yes)
    addlog(cool, out, norm, chrom, canny, minL, maxW, maxpixel, numcores, pvalue, mask, bfilter)
    print('Result will be stored in %s' % (out))
    chroms = chrom.split(',')
    cm = open_contact_map(cool, None if chroms[0] == 'all' else None)
    result_table, res_filter = run_stages(cm, norm, chrom, canny, minL, maxW, maxpixel, numcores, pvalue,
                                          mask, slow, bfilter, seed, device_id=device_id, verbose=True)
    result_table.to_csv(out + 'result_unfiltered.tsv', sep="\t", header=True, index=False)
    res_filter.to_csv(out + 'result_filtered.tsv', sep="\t", header=True, index=False)
    print('\n' + str(round((time.time() - t_start) / 60, 3)) + 'min taken.')
    print('Check the result stored in %s' % (out))
    return 0
