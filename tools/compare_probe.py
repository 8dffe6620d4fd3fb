"""Compare the partition-independent solution fingerprints (`config.solution_probe`) of two bench.py lines, e.g. config C
on 2 and on 4 GPUs at --rtol 1e-10:  python tools/compare_probe.py a.json b.json  -> one JSON line, rc 0 if <= 1e-8."""
import json
import sys


def probe(path):
    with open(path) as f:
        lines = [l for l in f.read().splitlines() if l.startswith("{")]
    d = json.loads(lines[-1])
    return d["n_gpus"], d["config"]["solution_probe"], d["config"]["cg_iterations"]


def main():
    na, a, ia = probe(sys.argv[1])
    nb, b, ib = probe(sys.argv[2])
    rel = lambda x, y: abs(x - y) / max(abs(x), abs(y), 1e-300)  # noqa: E731
    diffs = {"sum_u": rel(a["sum_u"], b["sum_u"]), "norm2_u": rel(a["norm2_u"], b["norm2_u"]),
             "max_abs_u": rel(a["max_abs_u"], b["max_abs_u"]),
             # moments that vanish by symmetry are compared on the scale of the largest one
             "weighted_moments": max(abs(x - y) for x, y in zip(a["weighted_moments"], b["weighted_moments"]))
             / max(abs(v) for v in a["weighted_moments"])}
    worst = max(diffs.values())
    out = {"gpus": [na, nb], "rtol": [a["rtol"], b["rtol"]], "iterations": [ia, ib], "relative_differences": diffs,
           "worst": worst, "bar": 1e-8, "ok": worst <= 1e-8}
    print(json.dumps(out))
    sys.exit(0 if out["ok"] else 1)


if __name__ == "__main__":
    main()
