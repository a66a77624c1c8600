"""squarna_b200 -- SQUARNA's greedy single-sequence hot path on B200 (sm_100a).

Same layout as the reference package: the modules SQRNdbnseq, SQRNdbnali and
SQUARNA carry the reference's function names (SQRNdbnseq / RunSQRNdbnseq,
SQRNdbnali / RunSQRNdbnali / YieldStems, Predict / Main / ParseConfig); the
package itself exports Predict and Main like the reference's __init__.py.
The compute path is libsqrn_b200.so (csrc/); there is no CPU fallback.
"""


def __getattr__(name):
    # lazy: importing the package must not need the CLI layer (or a GPU)
    if name in ("Predict", "Main", "ParseConfig"):
        from . import SQUARNA as _cli
        return getattr(_cli, name)
    raise AttributeError(name)
