"""``deepImpute(**kwargs)`` -- the reference's one-call entry point and console script (``deepImpute.py:6-40``).

Behaviour kept from the reference: the command line is always parsed first (``deepImpute.py:8``) and keyword
arguments then override the parsed values (``:10-11``); the CSV is read with the first column as index (``:13``)
and transposed for ``cell_axis='columns'`` (``:15-16``); the network is one ``Dense(hidden, relu)`` +
``Dropout(rate)`` block (``:18-27``); the result is written to ``output`` or returned when ``output`` is None
(``:34-37``).  Only the engine underneath differs.
"""
import sys

import pandas as pd

from .multinet import MultiNet
from .parser import build_parser, parse_args


def _resolve_args(kwargs):
    """argv first, kwargs on top.  Unlike the reference, a Python caller who passes ``inputFile=`` does not need
    a usable ``sys.argv`` (under pytest/Jupyter the reference's unconditional parse aborts)."""
    if "inputFile" in kwargs:
        parser = build_parser()
        args, _ = parser.parse_known_args([str(kwargs["inputFile"])])
    else:
        args = parse_args()
    for key, value in kwargs.items():
        setattr(args, key, value)
    return args


def deepImpute(**kwargs):
    args = _resolve_args(kwargs)

    data = pd.read_csv(args.inputFile, index_col=0)
    if args.cell_axis == "columns":
        data = data.T

    net = MultiNet(
        learning_rate=args.learning_rate,
        batch_size=args.batch_size,
        max_epochs=args.max_epochs,
        ncores=args.cores,
        sub_outputdim=args.output_neurons,
        architecture=[
            {"type": "dense", "activation": "relu", "neurons": args.hidden_neurons},
            {"type": "dropout", "activation": "dropout", "rate": args.dropout_rate},
        ],
        math_mode=getattr(args, "math", None),
    )
    net.fit(data, NN_lim=args.limit, cell_subset=args.subset, minVMR=args.minVMR, n_pred=args.n_pred)
    imputed = net.predict(data, imputed_only=False, policy=args.policy)

    if args.output is None:
        return imputed
    imputed.to_csv(args.output)


def main():
    deepImpute()
    return 0


if __name__ == "__main__":
    sys.exit(main())
