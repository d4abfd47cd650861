"""Shared helpers for the parity tests: load a golden fixture into oracle-style parameter dicts."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
MODEL_CASES = ["tiny_fp64", "f200_fp32", "transe_proj_fp32", "adversarial_fp32", "transe_fp64"]


class Case:
    def __init__(self, name):
        z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
        self.z = z
        self.name = name
        (self.n, self.e, self.r, self.d_in, self.f, self.h, self.layers, self.b, self.k,
         proj, self.proj_layers) = [int(v) for v in z["meta"]]
        self.projection = bool(proj)
        self.scorer = str(z["scorer"])
        self.loss_type = str(z["loss_type"])
        self.weights = tuple(float(w) for w in z["weights"])
        self.dtype = torch.float64 if z["x0"].dtype == np.float64 else torch.float32

    def t(self, key):
        return torch.from_numpy(self.z[key])

    def layer_prefix(self, li):
        return "gat_layer." if self.layers == 1 else f"gat_layers.{li}."

    def layer_params(self, requires_grad=False):
        out = []
        for li in range(self.layers):
            p = "param/" + self.layer_prefix(li)
            W = [self.t(f"{p}proj.{h}.weight").clone().requires_grad_(requires_grad) for h in range(self.h)]
            A = [self.t(f"{p}attn_vec.{h}").clone().requires_grad_(requires_grad) for h in range(self.h)]
            beta = self.t(f"{p}rel_bias").clone().requires_grad_(requires_grad)
            out.append({"W": W, "A": A, "beta": beta})
        return out

    def proj_params(self, requires_grad=False):
        if not self.projection:
            return None
        keys = [k for k in self.z.files if k.startswith("param/projection.net")]
        if "param/projection.net.weight" in keys:
            return {"weights": [self.t("param/projection.net.weight").clone().requires_grad_(requires_grad)], "ln": []}
        idx = sorted({int(k.split(".")[2]) for k in keys})
        weights, ln = [], []
        for i in idx:
            if f"param/projection.net.{i}.bias" in keys:
                ln.append((self.t(f"param/projection.net.{i}.weight").clone().requires_grad_(requires_grad),
                           self.t(f"param/projection.net.{i}.bias").clone().requires_grad_(requires_grad)))
            else:
                weights.append(self.t(f"param/projection.net.{i}.weight").clone().requires_grad_(requires_grad))
        return {"weights": weights, "ln": ln, "idx": idx}

    def rel_emb(self, requires_grad=False):
        return self.t("param/scorer.rel_emb.weight").clone().requires_grad_(requires_grad)

    def edge_index(self):
        return torch.from_numpy(np.stack([self.z["src"], self.z["dst"]]))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    denom = max(np.abs(b).max(), 1e-30) if b.size else 1.0
    return float(np.abs(a - b).max() / denom) if b.size else 0.0
