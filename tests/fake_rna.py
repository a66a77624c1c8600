"""A stand-in for ViennaRNA's `RNA` module with the surface SQUARNA's BPMatrix uses
(/root/reference/src/SQUARNA/SQRNdbnseq.py:341-365): fold_compound, sc_add_SHAPE_deigan, pf, bpp, mfe,
exp_params_rescale.  ViennaRNA is not installed in the build container (nor on the GPU box), so the
bpp != 0 parameter sets of def.conf / 500.conf / 1000.conf / greedy.conf cannot be pinned against it.
What CAN be pinned is everything around it: tests/golden/make_golden.py puts this module in
sys.modules["RNA"] and runs the REAL reference, and the product gets the same module through
squarna_b200.SQRNdbnseq.set_rna_module -- so the additive / multiplicative weighting, the powers,
the rescale fall-back, the SHAPE hand-over and every parameter set downstream are compared against
the reference's own code.  The "probabilities" are a deterministic function of the sequence and the
SHAPE vector: a skewed pseudo-random value on about a third of the canonical cells.
"""
import zlib

import numpy as np

_CANON = {"AU", "UA", "GC", "CG", "GU", "UG"}


class fold_compound:
    def __init__(self, sequence):
        self.seq = sequence
        self.shape = None
        self.rescaled = False
        self.calls = []

    def sc_add_SHAPE_deigan(self, reactivities, m=1.8, b=-0.6):
        self.shape = [float(x) for x in reactivities]
        self.calls.append(("shape", round(m, 6), round(b, 6)))

    def pf(self):
        self.calls.append(("pf",))
        return ("." * len(self.seq), -1.0)

    def mfe(self):
        return ("." * len(self.seq), -1.0 - 0.01 * len(self.seq))

    def exp_params_rescale(self, mfe):
        self.rescaled = True

    def bpp(self):
        """(N + 1) x (N + 1), 1-based upper triangle, like fold_compound.bpp()"""
        n = len(self.seq)
        seed = zlib.crc32(self.seq.encode("latin-1", "replace"))
        out = np.zeros((n + 1, n + 1))
        # every 5th sequence underflows at first and needs the rescaled partition function (seq.py:357-360);
        # every 13th gives nothing at all (the score matrix then stays as it is)
        if seed % 13 == 0 or (seed % 5 == 0 and not self.rescaled):
            return tuple(map(tuple, out.tolist()))
        rng = np.random.default_rng(seed)
        val = rng.random((n, n)) ** 4
        on = rng.random((n, n)) < 0.35
        for i in range(n):
            for j in range(i + 4, n):
                if on[i, j] and self.seq[i] + self.seq[j] in _CANON:
                    p = val[i, j]
                    if self.shape is not None:
                        p = p / (1.0 + 0.5 * abs(self.shape[i]) + 0.5 * abs(self.shape[j]))
                    out[i + 1, j + 1] = p
        return tuple(map(tuple, out.tolist()))
