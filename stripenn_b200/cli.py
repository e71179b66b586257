"""Command line — same commands, flags and defaults as reference src/stripenn/cli.py:8-58
(`stripenn compute`, `stripenn seeimage`, `stripenn score`).
Additive flags: --gpu (device index) and --yes (answer the overwrite prompt)."""
from __future__ import annotations

import multiprocessing

import typer

from . import compute as _compute
from . import score as _score
from . import seeimage as _seeimage

app = typer.Typer()


@app.command('compute')
def execute(
    cool: str = typer.Option(..., "--cool", help="Path to cool file (.npz contact map or .cool/.mcool URI)"),
    out: str = typer.Option(..., "--out", "-o", help="Path to output directory"),
    norm: str = typer.Option('KR', "--norm", help="Normalization method. It should be one of the column name of Cooler.bin(). Check it with Cooler.bins().columns (e.g., KR, VC, VC_SQRT)"),
    chrom: str = typer.Option('all', "--chrom", "-k", help="Set of chromosomes. e.g., 'chr1,chr2,chr3', 'all' will generate stripes from all chromosomes"),
    canny: float = typer.Option(2.0, "--canny", "-c", help="Canny edge detection parameter."),
    minL: int = typer.Option(10, '-l', '--minL', help="Minimum length of stripe."),
    maxW: int = typer.Option(8, '-w', '--maxW', help="Maximum width of stripe."),
    maxpixel: str = typer.Option('0.95,0.96,0.97,0.98,0.99', '-m', '--maxpixel', help="Percentiles of the contact frequency data to saturate the image."),
    numcores: int = typer.Option(multiprocessing.cpu_count(), '-n', '--numcores', help='The number of cores will be used (kept for compatibility; the GPU does the work).'),
    pvalue: float = typer.Option(0.1, '-p', '--pvalue', help='P-value cutoff for stripe.'),
    mask: str = typer.Option('0', "--mask", help='Column coordinates to be masked. e.g., chr9:12345678-12345789'),
    slow: bool = typer.Option(False, '-s', help='Use if system memory is low.'),
    bfilter: int = typer.Option(3, '--bfilter', '-b', help="Mean filter size. should be an odd number"),
    seed: int = typer.Option(123456789, '--seed', help="Seed of the background sampler."),
    gpu: int = typer.Option(0, '--gpu', help="CUDA device index."),
    yes: bool = typer.Option(False, '--yes', help="Overwrite an existing output directory without asking."),
):
    """Finds stripe coordinates from 3D genomic data"""
    _compute.compute(cool, out, norm, chrom, canny, minL, maxW, maxpixel, numcores, pvalue, mask, slow, bfilter,
                     seed, device_id=gpu, assume_yes=yes)


@app.command('seeimage')
def seeimag(
    cool: str = typer.Option(..., "--cool", help="Path to cool file"),
    position: str = typer.Option(..., "--position", '-p', help="Genomic position (e.g., chr1:135010000-136000000)"),
    maxpixel: str = typer.Option('0.95,0.96,0.97,0.98,0.99', "--maxpixel", '-m', help="Quantile for the pixel saturation. (e.g., 0.95)"),
    out: str = typer.Option('./heatmap.png', "--out", "-o", help="Path to output directory"),
    norm: str = typer.Option('KR', "--norm", help="Normalization method. It should be one of the column name of Cooler.bin(). Check it with Cooler.bins().columns (e.g., KR, VC, VC_SQRT)"),
    slow: bool = typer.Option(False, '-s', help='Use if system memory is low.'),
    seed: int = typer.Option(123456789, "--seed", help="Seed used to initialize the PRNG."),
    gpu: int = typer.Option(0, '--gpu', help="CUDA device index."),
):
    """ Draws heatmap image of given position and color saturation parameter (maxpixel)."""
    _seeimage.seeimage(cool, position, maxpixel, norm, out, slow, seed, device_id=gpu)
    return 0


@app.command('score')
def scoring(
    cool: str = typer.Option(..., "--cool", help="Path to cool file"),
    coordinates: str = typer.Option(..., "--coord", "-c", help="Path to stripe coordinate table"),
    norm: str = typer.Option('KR', "--norm", help="Normalization method."),
    numcores: int = typer.Option(multiprocessing.cpu_count(), '-n', '--numcores', help='The number of cores will be used.'),
    out: str = typer.Option('scores.out', "--out", "-o", help="Path to output file"),
    mask: str = typer.Option('0', "--mask", help='Column coordinates to be masked.'),
    seed: int = typer.Option(123456789, '--seed', help="Seed of the background sampler."),
    gpu: int = typer.Option(0, '--gpu', help="CUDA device index."),
):
    """Calculates p-value and stripiness of given stripes based on given 3D genome conformation data."""
    _score.getScore(cool, coordinates, norm, numcores, seed, out, mask, device_id=gpu)


def main():
    app()


if __name__ == "__main__":
    main()
