"""``python -m pycistopic_b200 topic_modeling lda ...``: the ``pycistopic topic_modeling lda`` command of the reference
(src/pycisTopic/cli/subcommand/topic_modeling.py:9-66 for the action, :171-302 for the flags) on top of
:func:`pycistopic_b200.run_cgs_models` -- same flags, same input (a pickled cisTopic object), same output (a pickled
list of ``CistopicLDAModel``; ``--keep`` also writes ``<output without .pkl>/Topic<K>.pkl``).  ``-p/--parallel`` is the
number of GPUs the models are spread over (the reference: Ray workers).  SURVEY.md section 8f, row N3; no other
subcommand of the reference CLI is rebuilt."""
from __future__ import annotations

import argparse
import pickle
import sys


def str_to_bool(v):
    """yes/true/t/y/1 and no/false/f/n/0, case-insensitive (topic_modeling.py:141-168)."""
    if isinstance(v, str):
        if v.lower() in ("yes", "true", "t", "y", "1"):
            return True
        if v.lower() in ("no", "false", "f", "n", "0"):
            return False
    raise argparse.ArgumentTypeError("Boolean value expected.")


def build_parser():
    parser = argparse.ArgumentParser(prog="pycistopic_b200", description="B200-native pycisTopic topic modelling")
    sub = parser.add_subparsers(dest="command", required=True)
    tm = sub.add_parser("topic_modeling", help="Run LDA topic modeling.")
    tm_sub = tm.add_subparsers(dest="topic_modeling", required=True)
    lda = tm_sub.add_parser("lda", help='"Run LDA topic modeling with "lda" package" -- here: the CUDA sampler.')
    lda.add_argument("-i", "--input", dest="input", type=str, required=True, help="cisTopic object pickle input filename.")
    lda.add_argument("-o", "--output", dest="output", type=str, required=True, help="Topic model list pickle output filename.")
    lda.add_argument("-t", "--topics", dest="topics", type=int, required=True, nargs="+",
                     help="Number(s) of topics to create during topic modeling.")
    lda.add_argument("-p", "--parallel", dest="parallel", type=int, required=True,
                     help="Number of topic models to run in parallel (= GPUs to use).")
    lda.add_argument("-n", "--iterations", dest="iterations", type=int, default=500)
    lda.add_argument("-a", "--alpha", dest="alpha", type=int, default=50)
    lda.add_argument("-A", "--alpha_by_topic", dest="alpha_by_topic", type=str_to_bool, default=True)
    lda.add_argument("-e", "--eta", dest="eta", type=float, default=0.1)
    lda.add_argument("-E", "--eta_by_topic", dest="eta_by_topic", type=str_to_bool, default=False)
    lda.add_argument("-k", "--keep", dest="keep_intermediate_topic_models", type=str_to_bool, default=False)
    lda.add_argument("-s", "--seed", dest="seed", type=int, default=555)
    lda.add_argument("-T", "--temp_dir", dest="temp_dir", type=str, default=None, help="Ignored (no Ray object store).")
    lda.set_defaults(func=run_topic_modeling_lda)
    return parser


def run_topic_modeling_lda(args):
    from . import _lib, compat, run_cgs_models
    compat.alias_pycistopic()   # reference pickles name pycisTopic.* classes; models written here carry those paths too
    out = args.output
    save_path = (out[:-4] if out.endswith(".pkl") else out) if args.keep_intermediate_topic_models else None
    print("Run topic modeling with the CUDA collapsed Gibbs sampler with the following settings:")
    for label, value in (("Input cisTopic object filename:", args.input), ("Topic modeling output filename:", out),
                         ("Number of topics to run topic modeling for:", args.topics), ("Alpha:", args.alpha),
                         ("Divide alpha by the number of topics:", args.alpha_by_topic), ("Eta:", args.eta),
                         ("Divide eta by the number of topics:", args.eta_by_topic),
                         ("Number of iterations:", args.iterations),
                         ("Number of GPUs the models are spread over:", args.parallel), ("Seed:", args.seed),
                         ("Save intermediate topic models in dir:", save_path)):
        print(f"  - {label:<45} {value}")
    print(f'\nLoading cisTopic object from "{args.input}"...\n')
    with open(args.input, "rb") as fh:
        cistopic_obj = pickle.load(fh)
    n_gpu = _lib.require_device()
    models = run_cgs_models(cistopic_obj, n_topics=args.topics, n_cpu=args.parallel, n_iter=args.iterations,
                            random_state=args.seed, alpha=args.alpha, alpha_by_topic=args.alpha_by_topic, eta=args.eta,
                            eta_by_topic=args.eta_by_topic, save_path=save_path, _temp_dir=args.temp_dir,
                            devices=list(range(max(1, min(n_gpu, args.parallel)))))
    print(f'\nWriting topic modeling output to "{out}"...')
    with open(out, "wb") as fh:
        pickle.dump(compat.pickle_ready(models), fh)
    return 0


def main(argv=None):
    args = build_parser().parse_args(argv)
    return args.func(args)


if __name__ == "__main__":
    sys.exit(main())
