"""Writes tests/golden/snapshot_ref.msgpack: a (small) snapshot produced by the REFERENCE's serializer -- nlohmann::json::to_msgpack over the document
Testbed::save_snapshot builds (oracle/ref_json.cpp: refjson_save_like_reference) from the reference's own configs/nerf/main.json.  Needs /root/reference
(for main.json) and `make -C oracle refjson`; run here, commit the output.  tests/test_snapshot_ref.py reads the file without either."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", ".."))
from oracle import pyrefjson  # noqa: E402

N, G = 4096, 128 ** 3 // 64  # small on purpose: the layout is what is pinned, not the size (G is not a whole cascade: the loader under test only stores it)


def state():
    r = np.random.default_rng(2024)
    return dict(params16=r.standard_normal(N).astype(np.float16), grid16=(r.random(G) * 0.05).astype(np.float16), m1=r.standard_normal(N).astype(np.float32),
                m2=r.random(N).astype(np.float32), steps=r.integers(0, 100, N).astype(np.uint32), ema16=r.standard_normal(N).astype(np.float16))


if __name__ == "__main__":
    text = open("/root/reference/configs/nerf/main.json").read()
    st = state()
    R = pyrefjson.RefJson()
    out = os.path.join(HERE, "snapshot_ref.msgpack")
    R.save_like_reference(out, text, st["params16"], st["grid16"], 4321, 0.03125, (0, 0, 0, 1, 1, 1), (0.25, 0.125, 0.375, 0.75, 0.875, 0.625), 1.0, 1, 81920, 262000, 3100000,
                          optimizer=dict(current_step=4321, base_lr=1.75e-3, m1=st["m1"], m2=st["m2"], steps=st["steps"], ema16=st["ema16"]))
    print(out, os.path.getsize(out), "bytes")
