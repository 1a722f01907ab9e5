"""Loads tests/golden/range_coder_golden.npz (generated from the compiled reference by oracle/make_golden.py)."""
import os

import numpy as np

PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "range_coder_golden.npz")


def load():
  return dict(np.load(PATH))


def split(flat, lens):
  out, at = [], 0
  for n in lens:
    out.append(bytes(flat[at:at + int(n)]))
    at += int(n)
  return out
