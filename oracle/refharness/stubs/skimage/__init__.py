"""Stub package standing in for scikit-image (not installed, no network).
Only ``skimage.feature.canny`` is provided — the single symbol the reference uses
(reference: src/stripenn/getStripe.py:11,917).
TEST INFRASTRUCTURE ONLY (oracle/refharness) — never imported by the product."""
__version__ = "0.0-restated"
