#!/usr/bin/env python
"""Parses EVERYTHING the reference holds about the 32x32 intra predictor out of src/mkIntra32-wip.bsv into
tests/golden/intra_bsv.json (runs only where /root/reference exists; the committed JSON travels).

The BSV is work in progress and does not compile, so no prediction can be produced from it; what it does hold is data:
  * mapTbl / facTbl / mapShift with the mode labels of their row comments            (:75-132)
  * getRefPixels: for every case the exact list of reference samples that forms the working line, i.e. the
    inverse-angle projection lists of all 14 negative-angle modes and the main-reference extents         (:135-328)
  * the live interpolator expression and the disabled block's rounding                 (:352-368, :503)
  * the DC accumulation                                                                 (:388-392, :318, :512-516)
tests/test_oracle.py checks the restatement (oracle/x266_oracle.c) against each of these, with the WIP file's
defects listed as explicit expected differences.
"""
import json
import os
import re
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_TREE = os.environ.get("X266_REF", "/root/reference")


def line_of(src, pos):
    return src.count("\n", 0, pos) + 1


def parse_table(src, name):
    m = re.search(name + r"\[\d+\]\[\d+\]\s*=\s*\{(.*?)\n\s*\};", src, re.S)
    body = m.group(1)
    rows, labels = [], []
    for rm in re.finditer(r"\{([^{}]*)\}\s*,?\s*//([^\n]*)", body):
        rows.append([int(v) for v in re.findall(r"-?\d+", rm.group(1))])
        labels.append(rm.group(2).strip())
    return {"rows": rows, "labels": labels, "lines": [line_of(src, m.start()), line_of(src, m.end())]}


def parse_ref_lines(src):
    f0 = src.index("function Vector#(64, Bit#(8)) getRefPixels")
    f1 = src.index("endfunction", f0)
    body = src[f0:f1]
    out = []
    for cm in re.finditer(r"^\s*([\d,\s]+):\s*begin(.*?)\n\s*end\b", body, re.S | re.M):
        cases = [int(v) for v in re.findall(r"\d+", cm.group(1))]
        um = re.search(r"unpack\(\{\?,(.*?)\}\);", cm.group(2), re.S)
        toks = re.findall(r"x([LT])\[\s*(\d+)\]|\b(dc)\b", um.group(1))
        entries = [[t[0], int(t[1])] if t[0] else ["dc", 0] for t in toks]
        entries.reverse()                       # BSV bit concatenation lists the highest vector index first: y[0] is the LAST token
        out.append({"cases": cases, "entries": entries, "line": line_of(src, f0 + cm.start(1))})
    return out


def main():
    path = os.path.join(REF_TREE, "src", "mkIntra32-wip.bsv")
    src = open(path).read()
    live = re.search(r"Bit#\(13\) tmp = zeroExtend\(v\[0\]\[i\]\) \* fromInteger\((.*?)\) \+ zeroExtend\(v\[1\]\[i\]\) \* fromInteger\((.*?)\) \+ (\d+);\s*"
                     r"z\[i\] = truncate\(tmp >> (\d+)\);", src)
    dead = re.search(r"y\[i\] = roundN\(\(x\[0\]\[i\] \+ x\[1\]\[i\]\), (\d+)\);", src)
    dc1 = re.search(r"Bit#\(14\) sum = 0;\s*for\(Integer i = 0; i < (\d+); i = i \+ 1\) begin\s*"
                    r"sum = sum \+ zeroExtend\((xL\[1\+i\])\) \+ zeroExtend\((xT\[1\+i\])\);\s*end\s*dcVal <= truncate\(sum >> (\d+)\);", src)
    dc2 = re.search(r"Bit#\(8\) dc = truncate\(dcVal >> (\d+)\);", src)
    dc3 = re.search(r"sum = sum \+ zeroExtend\((x\.left\[i\])\) \+ zeroExtend\((x\.top\[1\+i\])\);\s*end\s*dcVal <= truncate\(sum >> (\d+)\);", src)
    refs = re.search(r"Vector#\((\d+), Bit#\(8\)\) left;\s*Vector#\((\d+), Bit#\(8\)\) top;", src)
    out = {
        "source": "chenm001/x266 @ 379268c src/mkIntra32-wip.bsv (parsed, not executed: the file does not compile)",
        "refs": {"left": int(refs.group(1)), "top": int(refs.group(2)), "line": line_of(src, refs.start())},
        "mapTbl": parse_table(src, "mapTbl"),
        "facTbl": parse_table(src, "facTbl"),
        "mapShift": parse_table(src, "mapShift"),
        "ref_lines": parse_ref_lines(src),
        "interp_live": {"w0": live.group(1).strip(), "w1": live.group(2).strip(), "round": int(live.group(3)), "shift": int(live.group(4)),
                        "line": line_of(src, live.start())},
        "interp_disabled_block": {"roundN": int(dead.group(1)), "line": line_of(src, dead.start())},
        "dc": {"count": int(dc1.group(1)), "terms": [dc1.group(2), dc1.group(3)], "shift": int(dc1.group(4)), "line": line_of(src, dc1.start()),
               "second_shift_in_getRefPixels": int(dc2.group(1)), "second_shift_line": line_of(src, dc2.start()),
               "mkIntra32_terms": [dc3.group(1), dc3.group(2)], "mkIntra32_shift": int(dc3.group(3)), "mkIntra32_line": line_of(src, dc3.start())},
    }
    json.dump(out, open(os.path.join(HERE, "intra_bsv.json"), "w"), indent=0)
    print("intra_bsv.json:", {k: (len(v["rows"]) if isinstance(v, dict) and "rows" in v else len(v) if isinstance(v, list) else "...") for k, v in out.items()})


if __name__ == "__main__":
    sys.exit(main())
