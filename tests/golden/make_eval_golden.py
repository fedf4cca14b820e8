"""Generates tests/golden/eval_golden.json from the reference's own Evaluation() / Verify()
(network_helper.cpp compiled unmodified into oracle/_ref/libtf2ref_host_<net>.so by oracle/build_ref.sh).
Run in the build container only (needs /root/reference):  python tests/golden/make_eval_golden.py"""
import json
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", ".."))
from tests.test_verify_eval import eval_cases, ref_evaluation, ref_verify_lines, verify_case  # noqa: E402

out = {"evaluation": {}, "verify": {}}
with tempfile.TemporaryDirectory() as tmp:
    os.chdir(tmp)
    for name, net, x, q in eval_cases():
        out["evaluation"][name] = ref_evaluation(net, x, q, image=1)
    for net in ("resnet50", "googlenet"):
        fmap, expect, q = verify_case(net)
        rows = ref_verify_lines(net, fmap, expect, q, image=2, tmp=tmp)
        err = np.float32(0)
        tot = np.float32(0)
        for r in rows:
            err = np.float32(err + np.float32(r["error"]))
            tot = np.float32(tot + np.float32(r["expect_trans"]))
        out["verify"][net] = {"n": len(rows), "sum_error": float(err), "sum_expect": float(np.float32(sum(abs(r["expect_trans"]) for r in rows))),
                              "first_addrs": [r["addr"] for r in rows[:40]], "last_addr": rows[-1]["addr"]}
path = os.path.join(os.path.dirname(os.path.abspath(__file__)), "eval_golden.json")
with open(path, "w") as f:
    json.dump(out, f, indent=1)
print("wrote", path)
