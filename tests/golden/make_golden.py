#!/usr/bin/env python3
"""Build tests/golden/s288c_fixture.npz from the reference's own test fixtures.

Run in the build container (where /root/reference is mounted); the .npz travels to the
GPU box, /root/reference does not.  Nothing here is reference *source*: these are the
data files the reference's tests read (tests/S288c/*) and the golden output they pin
(tests/S288c/I.peaks.tsv, produced by the README's `gams wave` command, README.md:155-170).

Contents
  chr_names          ["I", "Mito"]
  seq_I, seq_Mito    uint8 ASCII bases exactly as in genome.fa.gz (mixed case kept)
  peaks_tsv          bytes of tests/S288c/I.peaks.tsv        (golden #1)
  snp_rg             bytes of tests/S288c/SK1.snp.rg         (golden #5)
  spo11_rg           bytes of tests/S288c/spo11_hot.rg       (golden #6)
  ctg_tsv            bytes of tests/S288c/ctg.tsv            (golden #7)
  intergenic_json    bytes of tests/S288c/intergenic.json    (anno fixture, §8f-4)
  genome_fa_gz       bytes of tests/S288c/genome.fa.gz       (a gzip member as the reference's loaders read it; gunzip golden)
"""
import gzip
import hashlib
import pathlib
import sys

import numpy as np

REF = pathlib.Path(sys.argv[1] if len(sys.argv) > 1 else "/root/reference") / "tests" / "S288c"
OUT = pathlib.Path(__file__).with_name("s288c_fixture.npz")


def read_fasta(path):
    names, seqs, cur = [], [], []
    with gzip.open(path, "rb") as fh:
        for line in fh:
            line = line.rstrip(b"\r\n")
            if line.startswith(b">"):
                if names:
                    seqs.append(b"".join(cur))
                names.append(line[1:].split()[0].decode())
                cur = []
            else:
                cur.append(line)
    seqs.append(b"".join(cur))
    return names, seqs


def main():
    names, seqs = read_fasta(REF / "genome.fa.gz")
    sizes = dict(l.split() for l in (REF / "chr.sizes").read_text().splitlines())
    for n, s in zip(names, seqs):
        assert len(s) == int(sizes[n]), (n, len(s), sizes[n])
    peaks = (REF / "I.peaks.tsv").read_bytes()
    assert hashlib.md5(peaks).hexdigest() == "3f6878c8daf92c67bee0c72805110f66"
    arrs = {
        "chr_names": np.array(names),
        "peaks_tsv": np.frombuffer(peaks, dtype=np.uint8),
        "snp_rg": np.frombuffer((REF / "SK1.snp.rg").read_bytes(), dtype=np.uint8),
        "spo11_rg": np.frombuffer((REF / "spo11_hot.rg").read_bytes(), dtype=np.uint8),
        "ctg_tsv": np.frombuffer((REF / "ctg.tsv").read_bytes(), dtype=np.uint8),
        "intergenic_json": np.frombuffer((REF / "intergenic.json").read_bytes(), dtype=np.uint8),
        "genome_fa_gz": np.frombuffer((REF / "genome.fa.gz").read_bytes(), dtype=np.uint8),
    }
    for n, s in zip(names, seqs):
        arrs["seq_" + n] = np.frombuffer(s, dtype=np.uint8)
    np.savez_compressed(OUT, **arrs)
    print("wrote", OUT, OUT.stat().st_size, "bytes;", {n: len(s) for n, s in zip(names, seqs)})


if __name__ == "__main__":
    main()
