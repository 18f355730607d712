#!/usr/bin/env python
"""Regenerates tests/golden/* from the UNMODIFIED reference.  Runs only where /root/reference exists
(this container); the GPU box uses the committed outputs.

  kat.json            FNV-1a-64 hashes / first values of the SURVEY.md 8(c) known-answer vectors, minted by
                      running oracle/_ref (= src_tb/dct32.c + src_tb/satd.c compiled in place).
  ref_vectors.npz     small input/output sets produced by the reference itself (random 9/11/16-bit and
                      extreme blocks, the srand(1) BDPI stream of the testbench).
  intra_bsv.json      every table, reference-line / projection list, the interpolator and the DC sum parsed from
                      src/mkIntra32-wip.bsv:75-132,135-328,352-392 (gen_intra_golden.py).
"""
import json
import os
import re
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import Oracle, Ref  # noqa: E402

REF_TREE = os.environ.get("X266_REF", "/root/reference")


def main():
    o, r = Oracle(), Ref()
    kat = {}
    for name, kind, shifts in (("A", 0, (4, 11)), ("B", 1, (6, 11)), ("C", 2, (4, 11))):
        x = o.residual(2040 * 1024, 266, kind)
        y = r.dct32(x.reshape(-1, 32, 32), *shifts)
        kat["KAT-" + name] = dict(op="dct32", blocks=2040, seed=266, kind=kind, shifts=shifts,
                                  fnv_in=f"{o.fnv(x):016x}", fnv_out=f"{o.fnv(y):016x}",
                                  out_first4=[int(v) for v in y.ravel()[:4]])
    for name, kind in (("D", 0), ("E", 2)):
        x = o.residual(32400 * 64, 266, kind)
        y = r.satd(x)
        kat["KAT-" + name] = dict(op="satd8x8", blocks=32400, seed=266, kind=kind,
                                  fnv_in=f"{o.fnv(x):016x}", fnv_out=f"{o.fnv(y):016x}",
                                  out_first4=[int(v) for v in y[:4]])
    # config 5 flavour: 8K frame, 11-bit, shifts 6/11 (one frame = 32400 blocks)
    x = o.residual(32400 * 1024, 266, 1)
    y = r.dct32(x.reshape(-1, 32, 32), 6, 11, threads=8)
    kat["KAT-8K"] = dict(op="dct32", blocks=32400, seed=266, kind=1, shifts=(6, 11),
                         fnv_in=f"{o.fnv(x):016x}", fnv_out=f"{o.fnv(y):016x}",
                         out_first4=[int(v) for v in y.ravel()[:4]])
    # the testbench's own first block: glibc srand(1)
    r.srand(1)
    diff, words, mat, dct = r.bdpi_dct_block()
    r.srand(1)
    rows, satd = r.bdpi_satd_block()
    kat["srand1"] = dict(dct_in_first4=[int(v) for v in mat[:4]], dct_out_first4=[int(v) for v in dct[:4]],
                         dct_fnv_out=f"{o.fnv(dct):016x}", getDiff_word0=f"{int(diff[0][0]):08x}",
                         getDiff_word16=f"{int(diff[0][16]):08x}", getDct_word0=f"{int(words[0]):016x}",
                         satd_first=satd)
    json.dump(kat, open(os.path.join(HERE, "kat.json"), "w"), indent=1)

    # small reference-produced vectors
    rng = np.random.default_rng(266)
    blocks = [o.residual(4 * 1024, 1, 0), o.residual(4 * 1024, 2, 1), o.residual(4 * 1024, 3, 2)]
    for v in (255, -255, 1023, -1023, 32767, -32768):
        blocks.append(np.full(1024, v, np.int16))
    alt = np.where((np.arange(1024) % 2) == 0, 32767, -32768).astype(np.int16)
    chk = np.where(((np.arange(1024) // 32 + np.arange(1024)) % 2) == 0, 32767, -32768).astype(np.int16)
    blocks += [alt, chk]
    for k in (0, 1, 17, 31):        # impulses read back columns of g_t32
        imp = np.zeros(1024, np.int16)
        imp[k] = 256
        blocks.append(imp)
    dct_in = np.concatenate(blocks).reshape(-1, 32, 32)
    sat_in = np.concatenate([o.residual(32 * 64, 4, 0), o.residual(32 * 64, 5, 2),
                             np.full(64, 255, np.int16), np.full(64, 1023, np.int16), np.full(64, -32768, np.int16),
                             np.where(np.arange(64) % 2 == 0, 255, -255).astype(np.int16),
                             rng.integers(-32768, 32768, 16 * 64).astype(np.int16)]).reshape(-1, 64)
    line_src = o.residual(5 * 32, 6, 2)
    np.savez_compressed(
        os.path.join(HERE, "ref_vectors.npz"),
        dct_in=dct_in, dct_out_4_11=r.dct32(dct_in, 4, 11), dct_out_6_11=r.dct32(dct_in, 6, 11),
        dct_out_1_1=r.dct32(dct_in, 1, 1), dct_out_9_16=r.dct32(dct_in, 9, 16),
        satd_in=sat_in, satd_out=r.satd(sat_in),
        partial_src=line_src, partial_out_shift4_line5=r.partial32(line_src, 4, 5),
        srand1_getDiff=diff, srand1_getDct=words, srand1_satd_rows=rows)

    # everything the BSV holds about the intra predictor (tables, projection lists, interpolator, DC): intra_bsv.json
    import gen_intra_golden
    gen_intra_golden.main()
    # the reference's own SAD golden dataset (riscv/programs/benchmarks/sad/dataset1.h: two 64x64 inputs + verify_data)
    ds = open(os.path.join(REF_TREE, "riscv", "programs", "benchmarks", "sad", "dataset1.h")).read()
    arrs = re.findall(r"(input_data1|input_data2|verify_data)\[DATA_SIZE\]\s*=\s*\{(.*?)\};", ds, re.S)
    got = {name: np.array([int(v) for v in re.findall(r"-?\d+", body)]) for name, body in arrs}
    np.savez_compressed(os.path.join(HERE, "sad_dataset.npz"), a=got["input_data1"].astype(np.uint8),
                        b=got["input_data2"].astype(np.uint8), verify=got["verify_data"][:1].astype(np.int64))
    print("golden written:", sorted(os.listdir(HERE)))


if __name__ == "__main__":
    main()
