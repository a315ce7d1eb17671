"""Whole-run cases for the design-loop host (pcramp_b200/host/pcramp_b200 vs the stock program oracle/_ref/pcramp): synthetic FASTA
inputs written from pcramp_b200.synth and the command-line flags both programs receive."""
import os

import numpy as np

from pcramp_b200 import synth


class DesignCase:
    def __init__(self, name, targets, flags, backgrounds=None):
        self.name, self.targets, self.backgrounds, self.flags = name, targets, backgrounds, flags


def write_fasta(path, coll, prefix):
    with open(path, "w") as fh:
        for i in range(coll.n):
            s = coll.text(i)
            fh.write(">%s%d\n" % (prefix, i))
            for k in range(0, len(s), 70):          # line width 70, deflines >t<i> / >b<i>
                fh.write(s[k:k + 70] + "\n")


def sister(seed, coll_factory_args, away):
    """backgrounds drawn from a sister ancestor `away` from the targets' root"""
    n, length, within = coll_factory_args
    fac = synth.TargetFactory(seed, n, length, n_clades=1, between=away, within=within)
    return fac.collection()


def cases():
    out = []
    # BASELINE config 1: 100 x 10 kb at 3 %, --seed 42 --thread 1, three assays so that splits, pool and multiplex database engage
    out.append(DesignCase("c1_seed42_count3", synth.make_targets(1, 100, 10000, within=0.03), ["--seed", "42", "--count", "3"]))
    # all six moves
    out.append(DesignCase("c1_moves", synth.make_targets(1, 100, 10000, within=0.03),
                          ["--seed", "7", "--count", "3", "--trial", "150", "-d", "4", "--optimize.5", "--optimize.3"]))
    # backgrounds: a sister clade 10 % away (the targets' root mutated), thresholds at their defaults (0.8 x 0.9)
    tf = synth.TargetFactory(2, 60, 4000, n_clades=1, between=0.0, within=0.02)
    bf = synth.TargetFactory(2, 20, 4000, n_clades=1, between=0.10, within=0.02)
    out.append(DesignCase("background", tf.collection(), ["--seed", "11", "--count", "3", "--trial", "200"], backgrounds=bf.collection()))
    # run to exhaustion: every target detected -> the active flags are reset and the major assay id advances (main.cpp:490-502)
    out.append(DesignCase("exhaust", synth.make_targets(9, 12, 3000, n_clades=2, between=0.2, within=0.01),
                          ["--seed", "5", "--count", "8", "--trial", "100"]))
    return out


def materialise(case, directory):
    """-> argv tail shared by both programs (without -o)"""
    t = os.path.join(directory, case.name + "_t.fa")
    write_fasta(t, case.targets, "t")
    argv = ["-t", t]
    if case.backgrounds is not None:
        b = os.path.join(directory, case.name + "_b.fa")
        write_fasta(b, case.backgrounds, "b")
        argv += ["-b", b]
    return argv + ["--thread", "1"] + list(case.flags)


def report_lines(path):
    """the report without its first two lines (program banner, command line)"""
    with open(path) as fh:
        lines = fh.read().splitlines()
    return [x for x in lines if not x.startswith("Command line:") and not x.startswith("PCRamp version")]
