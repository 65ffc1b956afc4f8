#!/usr/bin/env python
"""Extract the reference's own golden vectors into tests/golden/reference_vectors.json.

Run in the build container only (reads /root/reference, which does not exist on the GPU box):
    python tests/golden/make_golden.py
The JSON it writes is committed; tests never read /root/reference.

Everything here is DATA the reference's tests assert on (file:line cited per entry); no reference
code is executed (Julia is not installed) or copied.
"""
import json
import os
import re
import xml.etree.ElementTree as ET

REF = "/root/reference"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_vectors.json")


def true_pot():
    src = open(os.path.join(REF, "test/test_factors.jl")).read()
    m = re.search(r"true_pot\s*=\s*\[(.*?)\]", src, re.S)
    vals = [float(x) for x in re.findall(r"[-+]?\d+\.\d+", m.group(1))]
    assert len(vals) == 324
    return vals


def xdsl():
    root = ET.parse(os.path.join(REF, "test/sample_bn.xdsl")).getroot()
    out = []
    for cpt in root.find("nodes"):
        out.append({
            "id": cpt.get("id"),
            "states": [s.get("id") for s in cpt.findall("state")],
            "parents": (cpt.find("parents").text.split() if cpt.find("parents") is not None else []),
            "probabilities": [float(x) for x in cpt.find("probabilities").text.split()],
        })
    return out


def main():
    g = {
        "_about": "golden vectors asserted by the reference's own tests (sisl/BayesNets.jl v3.5.2); see make_golden.py",
        "factor_product": {
            "cite": "test/test_factors.jl:162-199",
            "phi1": {"dims": ["X", "C", "Y", "A", "Z"], "card": [3, 2, 3, 2, 3], "potential": "1:108"},
            "phi2": {"dims": ["A", "B", "C"], "card": [2, 3, 2], "potential": "1:12"},
            "true_pot": true_pot(),
        },
        "factor_product_mismatch": {
            "cite": "test/test_factors.jl:202-207",
            "phi1": {"dims": ["X", "Y", "Z"], "card": [3, 3, 3]},
            "phi2": {"dims": ["A", "X", "Y"], "card": [2, 31, 3]},
            "throws": "DimensionMismatch",
        },
        "reducedim": {
            "cite": "test/test_factors.jl:111-130",
            "dims": ["X", "Y", "Z"], "card": [3, 2, 2],
            "potential": [1, 2, 3, 2, 3, 4, 4, 6, 7, 8, 10, 16],
            "broadcast_mul_Z": [1, 10.0],
            "sum_over": ["Y", "Z"],
            "expected": [123.0, 165.0, 237.0],
            "zero_case": {"broadcast_mul_Z": 0.0, "sum_over": ["X", "Y", "Z"], "expected": [0.0]},
        },
        "normalize": {
            "cite": "test/test_factors.jl:71-81",
            "potential_colmajor": [1, 3, 2, 4],  # Float64[1 2; 3 4]
            "card": [2, 2],
            "p1": [0.1, 0.3, 0.2, 0.4],
            "p2": [1 / 30, 1 / 10, 2 / 30, 4 / 30],
        },
        "getindex": {
            "cite": "test/test_factors.jl:138-154",
            "dims": ["X", "Y", "Z"], "card": [3, 2, 2],
            "potential": [1, 2, 3, 2, 3, 4, 4, 6, 7, 8, 10, 16],
            "assignment_1based": {"Y": 2, "K": 16, "Z": 1},
            "expected": [2.0, 3.0, 4.0],
        },
        "setindex": {
            "cite": "test/test_factors.jl:138-154 (same phi as getindex; :147-152 the two assignments)",
            "dims": ["X", "Y", "Z"], "card": [3, 2, 2],
            "potential": [1, 2, 3, 2, 3, 4, 4, 6, 7, 8, 10, 16],
            "steps": [
                {"assignment_1based": {"X": 2, "Y": 1, "Z": 2}, "value": 1600.0,
                 "check_index_1based": [2, 1, 2], "expected": 1600.0},
                {"assignment_1based": {"X": 1, "Y": 2}, "value": 2016,
                 "check_slab_1based": {"X": 1, "Y": 2}, "expected_slab": [2016.0, 2016.0]},
            ],
        },
        "inference_acb": {
            "cite": "test/test_inference.jl:36-74",
            "nodes": [["a", [], [[1.0, 0.0]]], ["b", [], [[0.0, 1.0]]],
                      ["c", ["a", "b"], [[0.1, 0.9], [0.2, 0.8], [1.0, 0.0], [0.4, 0.6]]]],
            "P_a": [1.0, 0.0], "P_c": [1.0, 0.0],
            "P_bc_colmajor": [0.0, 1.0, 0.0, 0.0],
            "atol": 0.02,
        },
        "inference_student": {
            "cite": "test/test_inference.jl:76-110",
            "nodes": [["D", [], [[0.6, 0.4]]], ["I", [], [[0.7, 0.3]]],
                      ["G", ["D", "I"], [[0.3, 0.4, 0.3], [0.9, 0.08, 0.02], [0.05, 0.25, 0.7], [0.5, 0.3, 0.2]]],
                      ["L", ["G"], [[0.1, 0.9], [0.4, 0.6], [0.99, 0.01]]],
                      ["S", ["I"], [[0.95, 0.05], [0.2, 0.8]]]],
            "P_D": [0.6, 0.4],
            "P_G_given_d1_i1": [0.3, 0.4, 0.3],
            "atol": 0.05,
        },
        "xdsl": {
            "cite": "test/sample_bn.xdsl:4-15; test/test_io.jl:7-18",
            "cpts": xdsl(),
            "joint_1based": [[1, 1, 0.2 * 0.4], [1, 2, 0.2 * 0.4], [1, 3, 0.2 * 0.2],
                             [2, 1, 0.8 * 0.1], [2, 2, 0.8 * 0.3], [2, 3, 0.8 * 0.6]],
        },
        "text_io": {
            "cite": "test/test_io.jl:21-37",
            "lines": ["Success Forecast Test", "010", "001", "000", "2 3 2",
                      "0.2 0.4 0.4 0.1 0.3 1 0 0.1234569"],
        },
        "asia": {
            "cite": "src/gen_bayes_nets.jl:124-158",
            "nodes": [["A", [], [[0.99, 0.01]]], ["S", [], [[0.5, 0.5]]],
                      ["T", ["A"], [[0.99, 0.01], [0.95, 0.05]]],
                      ["L", ["S"], [[0.99, 0.01], [0.9, 0.1]]],
                      ["B", ["S"], [[0.97, 0.03], [0.4, 0.6]]],
                      ["X", ["T", "L"], [[0.95, 0.05], [0.02, 0.98], [0.02, 0.98], [0.02, 0.98]]],
                      ["D", ["T", "L", "B"], [[0.9, 0.1], [0.2, 0.8], [0.3, 0.7], [0.1, 0.9],
                                              [0.3, 0.7], [0.1, 0.9], [0.3, 0.7], [0.1, 0.9]]]],
        },
        "statistics_ab": {
            "cite": "test/test_discrete_bayes_nets.jl:13-67 (bn a -> b; data; bayesian_score_component values)",
            "names": ["a", "b"], "edges": [["a", "b"]],
            "data": {"a": [1, 1, 1, 1, 2, 2, 2, 2], "b": [1, 2, 1, 2, 1, 1, 1, 2]},
            "count_a": [4, 4],
            "count_b_given_a_1based": {"1,1": 2, "2,1": 3, "1,2": 2, "2,2": 1},  # (a, b) -> count
            "score_1_uniform": -6.445719819385579, "score_1_uniform2": -6.13556489108174,
            "score_1_bdeu": -6.841859646909766, "score_1_bdeu2": -6.445719819385579,
            "score_2_uniform": -6.396929655216146,
        },
        "statistics_abc": {
            "cite": "test/test_learning.jl:40-49 (data, dag A->B, A->C, B->C), :87-96 (statistics matrices)",
            "names": ["A", "B", "C"], "edges": [["A", "B"], ["A", "C"], ["B", "C"]],
            "data": {"A": [1, 1, 1, 1, 2, 2, 2, 2, 3, 3, 3, 3], "B": [1, 1, 1, 2, 2, 2, 2, 1, 1, 2, 1, 1],
                     "C": [1, 1, 1, 2, 1, 1, 2, 1, 1, 2, 1, 1]},
            "stats_A": [[4], [4], [4]],
            "stats_C": [[3, 1, 3, 0, 2, 0], [0, 0, 0, 1, 1, 1]],
        },
        "philox4x32_10_kat": {
            "cite": "Random123 kat_vectors (Salmon et al. SC'11), philox4x32 10 rounds",
            "vectors": [
                {"ctr": [0, 0, 0, 0], "key": [0, 0],
                 "out": [0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8]},
                {"ctr": [0xffffffff] * 4, "key": [0xffffffff] * 2,
                 "out": [0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd]},
                {"ctr": [0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344], "key": [0xa4093822, 0x299f31d0],
                 "out": [0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1]},
            ],
        },
    }
    with open(OUT, "w") as f:
        json.dump(g, f, indent=1)
    print("wrote", OUT)


if __name__ == "__main__":
    main()
