"""Loader for tests/golden/*.npz (written by oracle/make_golden.py from the unmodified reference)."""
import glob
import json
import os

import numpy as np
import torch

from literalkg_oracle import OracleConfig

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))


HEADS_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden_heads")
HEAD_CASES = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(HEADS_DIR, "*.npz")))


class Golden:
    def __init__(self, name, directory=GOLDEN_DIR):
        z = np.load(os.path.join(directory, name + ".npz"))
        self.name = name
        raw = json.loads(bytes(z["config"]).decode())
        self.n = raw.pop("n_entities")
        self.n_rel = raw.pop("n_relations")
        self.laplacian_type = raw.pop("laplacian_type")
        self.cfg = OracleConfig(**raw)
        self.sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
        self.z = {k: z[k] for k in z.files if not k.startswith("sd/") and k != "config"}

    def t(self, key):
        return torch.from_numpy(self.z[key])

    @property
    def num_lit(self):
        return self.t("in/num_lit") if self.cfg.use_num_lit else None

    @property
    def txt_lit(self):
        return self.t("in/txt_lit") if self.cfg.use_txt_lit else None

    def grads(self, mode):
        pre = f"grad_{mode}/"
        return {k[len(pre):]: torch.from_numpy(v) for k, v in self.z.items() if k.startswith(pre)}
