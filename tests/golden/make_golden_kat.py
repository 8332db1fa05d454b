#!/usr/bin/env python
"""Known-answer vectors that the REFERENCE'S OWN unit tests embed for the DE path
(anuga/shallow_water/tests/test_shallow_water_domain.py): the 8-digit expected arrays are lifted
from the test source into tests/golden/kat_reference_tests.npz (data only; the set-ups are restated
in cases.py: kat_bedslope_*).  usage: python tests/golden/make_golden_kat.py"""
import os
import re

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = "/root/reference/anuga/shallow_water/tests/test_shallow_water_domain.py"
TESTS = {
    "one_step": "test_bedslope_problem_second_order_one_step",
    "two_steps": "test_bedslope_problem_second_order_two_steps",
    "more_steps": "test_bedslope_problem_second_order_more_steps",
}


def body_of(text, name):
    i = text.index("def %s(" % name)
    j = text.index("\n    def test_", i + 10)
    return text[i:j], text[:i].count("\n") + 1


def array_after(body, label):
    """the LAST array assigned to `label` (the initial-condition check may use the same name)"""
    ms = re.findall(r"%s\s*=\s*(?:num\.array\()?\[(.*?)\]" % label, body, re.S)
    return None if not ms else np.array([float(x) for x in re.findall(r"-?\d+\.\d+(?:e-?\d+)?", ms[-1])])


def main():
    text = open(SRC).read()
    out = {}
    for key, name in TESTS.items():
        body, line = body_of(text, name)
        out[key + "_line"] = np.array([line])
        for label in ("W_EX", "UH_EX", "VH_EX"):
            a = array_after(body, label)
            if a is not None:
                out[key + "_" + label] = a
        m = re.search(r"evolve\(yieldstep\s*=\s*([\d.]+),\s*finaltime\s*=\s*([\d.]+)\)", body)
        out[key + "_evolve"] = np.array([float(m.group(1)), float(m.group(2))])
        for label in ("recorded_min_timestep", "recorded_max_timestep"):
            m = re.search(r"%s,\s*([\d.]+)\)" % label, body)
            if m:
                out[key + "_" + label] = np.array([float(m.group(1))])
    np.savez_compressed(os.path.join(HERE, "kat_reference_tests.npz"), **out)
    for k, v in sorted(out.items()):
        print(k, v.shape, v[:3])


if __name__ == "__main__":
    main()
