"""`chiron call` -- the command-line surface of chiron/entry.py:62-155 for the basecalling path.

    python -m chiron_b200.entry call -i <fast5 or signal folder> -o <out> [-m <model>] [-p dna-pre] [--beam 0] ...

Like the reference (entry.py:19-47) `call` first extracts every fast5 under the input folder to ``<out>/raw/*.signal``
(chiron/utils/extract_sig_ref.py) and then runs chiron_eval on that folder.  `export` and `train` are training-side
commands and are out of scope for this package."""
from __future__ import annotations

import argparse
import sys
from os import path

from . import __version__, chiron_eval
from .utils.extract_sig_ref import extract


def evaluation(args):
    args = chiron_eval.apply_preset(args)
    FLAGS = args
    FLAGS.input_dir = FLAGS.input
    FLAGS.output_dir = FLAGS.output
    FLAGS.unit = False
    FLAGS.recursive = True
    FLAGS.polya = None
    FLAGS.idname = False
    FLAGS.delimiter = "\n"
    extract(FLAGS)
    FLAGS.input = FLAGS.output + "/raw/"
    chiron_eval.run(args)


def main(arguments=None):
    arguments = sys.argv[1:] if arguments is None else arguments
    parser = argparse.ArgumentParser(prog="chiron", description="A deep neural network basecaller.")
    parser.add_argument("-v", "--version", action="version", version="chiron_b200 version " + __version__)
    subparsers = parser.add_subparsers(title="sub command", help="sub command help")
    parser_call = subparsers.add_parser("call", description="Perform basecalling", help="Perform basecalling.")
    chiron_eval.add_call_arguments(parser_call, model_default="DNA_default")
    parser_call.add_argument("--test_number", default=None, type=int, help="Extract test_number reads.")
    parser_call.set_defaults(func=evaluation)
    args = parser.parse_args(arguments)
    if hasattr(args, "func"):
        args.func(args)
    else:
        parser.print_help()


if __name__ == "__main__":
    main()
