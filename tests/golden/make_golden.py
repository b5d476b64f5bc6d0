"""Regenerate the committed golden fixtures from the reference checkout.

Run in the build container only (``/root/reference`` does not exist on the GPU box):

    python tests/golden/make_golden.py

Writes, next to this script:
  build_smoke.idx / build_smoke.idx.meta.json
      byte copies of the FAISS-written files the reference ships under
      pytest-tmp/build-smoke/data/ (IndexFlatIP, d=8, 1 row) -- the format known-answer.
  real77.index
      the nested ``IxFI`` block (77 x 4096 real embeddings) cut out of the FAISS ``IHNf``
      container data/photo_search.index, re-wrapped as a stand-alone flat index file.
  real77_hnsw_header.bin
      the first 30865 bytes of that container (HNSW graph) + nothing else, so the IHNf reader
      can be tested on the GPU box by concatenating it with real77.index.
  real77_time.json
      per-row ``exif_data.datetime`` and ``time_info`` of data/metadata.json (no paths, no text).
  real77_topk.json
      oracle top-10 of every stored row used as a query + predicate known-answers computed
      with the restated ``_check_time_match_v2``; a guard against oracle drift.
"""
from __future__ import annotations

import json
import os
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("PSX_REFERENCE", "/root/reference")

from oracle import flat_ip as O  # noqa: E402


def main() -> None:
    smoke = os.path.join(REF, "pytest-tmp", "build-smoke", "data")
    shutil.copyfile(os.path.join(smoke, "idx"), os.path.join(HERE, "build_smoke.idx"))
    shutil.copyfile(os.path.join(smoke, "idx.meta.json"), os.path.join(HERE, "build_smoke.idx.meta.json"))

    src = os.path.join(REF, "data", "photo_search.index")
    raw = open(src, "rb").read()
    index, info = O.read_index(src)
    off = info["storage_offset"]
    with open(os.path.join(HERE, "real77_hnsw_header.bin"), "wb") as f:
        f.write(raw[:off])
    with open(os.path.join(HERE, "real77.index"), "wb") as f:
        f.write(raw[off:])
    # the cut must itself be a valid flat file with identical rows
    again, _ = O.read_index(os.path.join(HERE, "real77.index"))
    assert np.array_equal(again._matrix(), index._matrix())

    meta = json.load(open(os.path.join(REF, "data", "metadata.json"), encoding="utf-8"))
    slim = [{"exif_data": {"datetime": (m.get("exif_data") or {}).get("datetime")}, "time_info": m.get("time_info")} for m in meta]
    json.dump(slim, open(os.path.join(HERE, "real77_time.json"), "w", encoding="utf-8"), ensure_ascii=False, indent=1)

    x = index._matrix()
    D, I = index.search(x, 10)
    constraints = [
        {"season": "夏天"},
        {"season": "冬天", "time_period": "下午"},
        {"year": 2023},
        {"year": 2023, "month": 8},
        {"start_date": "2023-01-01", "end_date": "2023-12-31"},
        {"start_date": "2024-06-01"},
        {"end_date": "2022-12-31T12:00:00"},
        {"season": "秋天", "start_date": "2021-01-01", "end_date": "2025-12-31"},
    ]
    preds = []
    for c in constraints:
        mask = [O.check_time_match_v2(m, c) for m in slim]
        Dm, Im = index.search(x[:4], 10, mask=np.array(mask))
        preds.append({"constraints": c, "pass_rows": [i for i, ok in enumerate(mask) if ok], "ids": Im.tolist(),
                      "scores": [[float(v) for v in row] for row in Dm]})
    json.dump(
        {"k": 10, "ids": I.tolist(), "scores": [[float(v) for v in row] for row in D], "predicates": preds,
         "ihnf": {k: info[k] for k in ("entry_point", "max_level", "efConstruction", "efSearch", "storage_offset")}},
        open(os.path.join(HERE, "real77_topk.json"), "w", encoding="utf-8"), ensure_ascii=False,
    )
    print("golden fixtures written to", HERE)


if __name__ == "__main__":
    main()
