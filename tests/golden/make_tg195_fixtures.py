"""Writes the AAPM TG-195 input fixtures the GPU box needs (it has no /root/reference):

  tests/golden/tg195_spectra.npz   the three tabulated TG-195 x-ray spectra the reference's validation program uses
                                   (validation/validation.cpp:148 100 kVp, :159 120 kVp, :172 30 kVp), as (energy, weight)
                                   with the reference's bin shift energy - 0.25 keV applied (TG195_specter, :175-186)
  tests/golden/case5world.tar.gz   byte copy of validation/data/case5world.tar.gz: the TG-195 Case 5 voxel phantom,
                                   500 x 320 x 260 material indices (u8), which the validation program reads as case5world.bin
                                   (validation.cpp:1290-1306, 1341)

    python tests/golden/make_tg195_fixtures.py [/root/reference]

Published data (AAPM TG-195 report), no reference code. Run once in the build container; the outputs are committed."""
import os
import re
import shutil
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF = sys.argv[1] if len(sys.argv) > 1 else "/root/reference"

src = open(os.path.join(REF, "validation", "validation.cpp")).read()
out = {}
for name in ("TG195_100KV_raw", "TG195_120KV_raw", "TG195_30KV_raw"):
    m = re.search(name + r"\(\{([^}]*)\}\)", src)
    raw = np.array([float(x) for x in m.group(1).split(",")], np.float64)
    key = name[len("TG195_"):-len("_raw")].lower()
    out[key + "_energy"] = (raw[0::2] - 0.25).astype(np.float32)
    out[key + "_weight"] = raw[1::2].astype(np.float32)
    print(name, raw.size // 2, "bins, mean energy", float((raw[0::2] * raw[1::2]).sum() / raw[1::2].sum()))
np.savez_compressed(os.path.join(HERE, "tg195_spectra.npz"), **out)
shutil.copyfile(os.path.join(REF, "validation", "data", "case5world.tar.gz"), os.path.join(HERE, "case5world.tar.gz"))
print("wrote tg195_spectra.npz and case5world.tar.gz")
