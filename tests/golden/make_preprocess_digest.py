#!/usr/bin/env python
"""Writes tests/golden/preprocess_digest.json: shape / statistics / 512 sampled voxels of the UNMODIFIED reference's
CTReportDataset.nii_img_to_tensor (scripts/data.py:92-162) on the seeded case of tests/test_oracle_cpu.py::_preprocess_case.
Run in the build container (needs /root/reference): python tests/golden/make_preprocess_digest.py"""
import json
import sys
from pathlib import Path

import pandas as pd

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import ref_shims  # noqa: E402
from tests.test_oracle_cpu import _digest, _preprocess_case  # noqa: E402

mod = ref_shims.load_reference_dataset_class()
raw, m = _preprocess_case()


class _Img:
    def get_fdata(self):
        return raw.astype("float64")


mod.nib.load = lambda path: _Img()
ds = object.__new__(mod.CTReportDataset)
df = pd.DataFrame({"VolumeName": ["case.nii.gz"], "RescaleSlope": [m["slope"]], "RescaleIntercept": [m["intercept"]],
                   "XYSpacing": [f"[{m['xy']}, {m['xy']}]"], "ZSpacing": [m["z"]]})
out = ds.nii_img_to_tensor("/data/case.nii.gz", df)
d = _digest(out)
(Path(__file__).parent / "preprocess_digest.json").write_text(json.dumps(d, indent=1) + "\n")
print(d)
