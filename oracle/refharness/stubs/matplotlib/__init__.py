"""Stub: the reference imports matplotlib only to call ``matplotlib.use('pdf')``
(reference: src/stripenn/getStripe.py:4,9).  No arithmetic lives here.
TEST INFRASTRUCTURE ONLY (oracle/refharness) — never imported by the product."""


def use(*_args, **_kwargs):
    return None
