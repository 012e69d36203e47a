"""Drop-in replacement of the reference's `transformer` package (same classes, constructor
signatures, forward signatures and state-dict schema), backed by hand-written sm_100a kernels.
Put this directory's parent on PYTHONPATH ahead of the reference checkout and the reference's
train.py / eval.py / synthesize.py import it unchanged."""
