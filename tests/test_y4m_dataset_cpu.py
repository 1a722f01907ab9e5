"""CPU: Y4MDataset against the reference's own test files (python/datasets/y4m_dataset_test.py:27-73) and the
reader's error classes (cc/kernels/y4m_dataset_kernels.cc:300-400)."""
import pytest
import torch

import compression_b200 as tfc


def _u8(string, shape):
  return torch.tensor(list(string), dtype=torch.uint8).reshape(shape)


def test_dataset_yields_correct_sequence(tmp_path):
  a, b = tmp_path / "a.y4m", tmp_path / "b.y4m"
  a.write_bytes(b"YUV4MPEG2 W4 H2 F30:1 Ip A0:0 C420jpeg\nFRAME\nABCDEFGHIJKL")
  b.write_bytes(b"YUV4MPEG2 C444 W1 H1\nFRAME\nabcFRAME\ndef")
  it = iter(tfc.Y4MDataset([str(a), str(b)]))
  y, cbcr = next(it)
  assert y.dtype == torch.uint8 and cbcr.dtype == torch.uint8
  assert torch.equal(y, _u8(b"ABCDEFGH", (2, 4, 1)))
  assert torch.equal(cbcr[..., 0], _u8(b"IJ", (1, 2))) and torch.equal(cbcr[..., 1], _u8(b"KL", (1, 2)))
  y, cbcr = next(it)
  assert torch.equal(y, _u8(b"a", (1, 1, 1))) and torch.equal(cbcr, _u8(b"bc", (1, 1, 2)))
  y, cbcr = next(it)
  assert torch.equal(y, _u8(b"d", (1, 1, 1))) and torch.equal(cbcr, _u8(b"ef", (1, 1, 2)))
  with pytest.raises(StopIteration):
    next(it)
  assert len(list(tfc.Y4MDataset(str(b)))) == 2          # a single file name, iterated afresh


@pytest.mark.parametrize("content,message", [
    (b"YUV4MPEG W2 H2 C444\n", "YUV4MPEG2 marker"),
    (b"YUV4MPEG2 W2 H2 C444", "complete Y4M header"),
    (b"YUV4MPEG2 W0 H2 C444\n", "invalid width"),
    (b"YUV4MPEG2 W2 Hx C444\n", "invalid height"),
    (b"YUV4MPEG2 W2 H2 C422\n", "unsupported chroma format '422'"),
    (b"YUV4MPEG2 W2 H2 It C444\n", "not in progressive"),
    (b"YUV4MPEG2 H2 C444\n", "no width"),
    (b"YUV4MPEG2 W2 C444\n", "no height"),
    (b"YUV4MPEG2 W2 H2\n", "no chroma format"),
    (b"YUV4MPEG2 W3 H2 C420jpeg\n", "odd width or height"),
    (b"YUV4MPEG2 W2 H2 C444\nFRAME\n0123", "incomplete or unsupported frame at byte 21"),
    (b"YUV4MPEG2 W1 H1 C444\nFRAMEx123", "FRAME marker at byte 21"),
    (b"YUV4MPEG2W1 H1 C444\n", "invalid Y4M header"),
])
def test_reader_errors(tmp_path, content, message):
  f = tmp_path / "bad.y4m"
  f.write_bytes(content)
  with pytest.raises(tfc.InvalidArgumentError, match=message):
    list(tfc.Y4MDataset([str(f)]))
