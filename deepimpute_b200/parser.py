"""Command-line flags of the ``deepImpute`` console script.

Same flag names, types and defaults as the reference's argparse set-up (reference ``deepimpute/parser.py:3-95``),
including its defaults that disagree with its own help strings (``--learning-rate`` 0.0005, ``--max-epochs`` 300,
``--hidden-neurons`` 300) -- a user switching over gets the same behaviour for the same command line.  Written as
a table so the whole surface is visible at a glance; one GPU-only flag (``--math``) is appended.
"""
import argparse

# (flags, kwargs) in the reference's order
_FLAGS = [
    (("inputFile",), dict(type=str, help="Path to input data.")),
    (("-o", "--output"), dict(type=str, default="./imputed.csv",
                              help="Path to output data counts. Default: ./imputed.csv")),
    (("--cores",), dict(type=int, default=-1,
                        help="Number of cores (accepted for compatibility; the GPU engine ignores it).")),
    (("--cell-axis",), dict(type=str, choices=["rows", "columns"], default="rows",
                            help="Cell dimension in the matrix. Default: rows")),
    (("--limit",), dict(type=str, default="auto", help="Genes to impute (e.g. first 2000 genes). Default: auto")),
    (("--minVMR",), dict(type=float, default=0.5,
                         help="Min variance over mean ratio for gene exclusion, used when --limit is 'auto'. "
                              "Default: 0.5")),
    (("--subset",), dict(type=float, default=1,
                         help="Cell subset used for training: a ratio (0<x<1) or a cell count. Default: 1 (all)")),
    (("--learning-rate",), dict(type=float, default=0.0005, help="Learning rate. Default: 0.0005")),
    (("--batch-size",), dict(type=int, default=64, help="Batch size. Default: 64")),
    (("--max-epochs",), dict(type=int, default=300, help="Maximum number of epochs. Default: 300")),
    (("--hidden-neurons",), dict(type=int, default=300,
                                 help="Number of neurons in the hidden dense layer. Default: 300")),
    (("--dropout-rate",), dict(type=float, default=0.2, help="Dropout rate of the hidden layer. Default: 0.2")),
    (("--output-neurons",), dict(type=int, default=512,
                                 help="Number of output neurons per sub-network. Default: 512")),
    (("--n_pred",), dict(type=int, default=None,
                         help="Number of candidate predictor genes. Default: all genes with nonzero VMR")),
    (("--policy",), dict(type=str, default="restore",
                         help="'restore' keeps every positive raw value, 'max' keeps max(raw, imputed). "
                              "Default: restore")),
    # ---- not in the reference ----
    (("--math",), dict(type=str, default=None, choices=["tf32x3", "tf32", "fp32"],
                       help="Arithmetic of the GPU engine: tf32x3 = tensor cores with error-compensated forward "
                            "products (default), tf32 = single-pass tensor cores, fp32 = CUDA cores.")),
]


def build_parser():
    parser = argparse.ArgumentParser(description="scRNA-seq data imputation using DeepImpute (B200 engine).")
    for flags, kw in _FLAGS:
        parser.add_argument(*flags, **kw)
    return parser


def parse_args(argv=None):
    """Parse ``sys.argv`` (or ``argv``).  The reference calls this even from Python (``deepImpute.py:8``)."""
    return build_parser().parse_args(argv)
