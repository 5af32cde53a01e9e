"""Writes the synthetic inputs of a BASELINE config as TSV for bench/reference_nullranges.R.
usage: python bench/export_inputs.py <cfg 1|2|3> <outdir>"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from nullranges_b200 import synth

cfg, out = int(sys.argv[1]), sys.argv[2]
os.makedirs(out, exist_ok=True)


def write(gr, name, extra=()):
    cols = [gr.seqnames, gr.start, gr.end] + [gr.mcols[k] for k in extra]
    with open(os.path.join(out, name + ".tsv"), "w") as f:
        f.write("\t".join(["seqnames", "start", "end", *extra]) + "\n")
        for row in zip(*cols):
            f.write("\t".join(str(v) for v in row) + "\n")


y = synth.uniform_ranges(10_000, seed=1) if cfg == 1 else synth.atac_peaks(1_000_000, seed=3)
write(y, "y")
write(synth.genes(20_000, seed=2), "x")
if cfg == 3:
    write(synth.segmentation(seed=4), "seg", extra=("state",))
    write(synth.exclude_regions(seed=5), "exclude")
with open(os.path.join(out, "seqlengths.tsv"), "w") as f:
    for k, v in synth.HG38.items():
        f.write(f"{k}\t{v}\n")
print("wrote", out)
