"""Writes tests/golden/npmath_digests.json and npmath_vectors.npz from THIS host's numpy.

Run in the build container (numpy 2.3.5, AVX-512 dispatch -- the build every golden in tests/golden/ comes from):
    python tests/make_npmath_golden.py
"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(__file__))
import npmath_vectors as V  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    feats = np.__config__.show_config(mode="dicts")["SIMD Extensions"]["found"] if hasattr(np.__config__, "show_config") else []
    out = {"numpy": np.__version__, "simd_found": feats, "n_per_set": V.N_PER_SET, "tanh": {}, "arctanh": {}}
    with np.errstate(all="ignore"):
        for name, x in V.tanh_sets().items():
            out["tanh"][name] = {"n": int(x.size), "sha256": V.digest(np.tanh(x))}
        for name, x in V.arctanh_sets().items():
            out["arctanh"][name] = {"n": int(x.size), "sha256": V.digest(np.arctanh(x))}
        cx = np.array([0.7310585786300049, -0.33, 0.9999999, 1e-5, 0.123456789])
        out["canary"] = {"tanh": np.tanh(cx).view(np.uint64).tolist(), "arctanh": np.arctanh(cx).view(np.uint64).tolist()}
        # explicit vectors: 8192 arguments per function with numpy's outputs
        xt = np.concatenate([s[:2048] for s in V.tanh_sets(2048).values()])
        xa = np.concatenate([s[:1700] for s in V.arctanh_sets(1700).values()])
        np.savez_compressed(os.path.join(HERE, "golden", "npmath_vectors.npz"),
                            tanh_x=xt.view(np.uint64), tanh_y=np.tanh(xt).view(np.uint64),
                            arctanh_x=xa.view(np.uint64), arctanh_y=np.arctanh(xa).view(np.uint64))
    with open(os.path.join(HERE, "golden", "npmath_digests.json"), "w") as f:
        json.dump(out, f, indent=1)
    print(json.dumps({k: v for k, v in out.items() if k not in ("tanh", "arctanh")}))
    print(sum(v["n"] for v in out["tanh"].values()), "tanh arguments,", sum(v["n"] for v in out["arctanh"].values()), "arctanh arguments")


if __name__ == "__main__":
    main()
