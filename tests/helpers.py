"""Shared helpers for the test-suite."""
import os

import numpy as np

from genlm_backend_b200 import Token

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return np.load(os.path.join(GOLDEN, name + ".npz"))


def unflat(blob, lens):
    """Inverse of make_golden.flat: list of byte strings."""
    out, at = [], 0
    data = blob.tobytes()
    for n in lens.tolist():
        out.append(data[at:at + n])
        at += n
    return out


def tokens(byte_strings):
    return [Token(i, b) for i, b in enumerate(byte_strings)]


def rel_err(have, want):
    """Max relative error where want != 0, and max |have| where want == 0 (must be exactly 0 for masses)."""
    have = np.asarray(have, dtype=np.float64)
    want = np.asarray(want, dtype=np.float64)
    nz = want != 0
    r = np.abs(have[nz] - want[nz]) / np.abs(want[nz]) if nz.any() else np.zeros(1)
    z = np.abs(have[~nz]) if (~nz).any() else np.zeros(1)
    return float(r.max()), float(z.max())


class EOS:
    """A non-byte vocabulary item: iterating it yields itself (the reference's EndOfSequence pattern)."""

    def __iter__(self):
        return iter([self])

    def __repr__(self):
        return "EOS"
