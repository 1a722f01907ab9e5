"""Y4MDataset: Y'CbCr video frames from '.y4m' files, mirroring tensorflow_compression/python/datasets/
y4m_dataset.py:25-63 and the reader behind it (cc/kernels/y4m_dataset_kernels.cc:120-400).

Host-side input plumbing, not part of the hot path: an iterable `torch.utils.data.IterableDataset` that reads every
file sequentially and yields, per frame, the luma plane `(H, W, 1)` and the interleaved chroma planes `(Hc, Wc, 2)`
as uint8 tensors.  Supported, as in the reference: progressive `C420jpeg` / `C420` (Hc = H/2, Wc = W/2) and `C444`;
every other header field is ignored; anything else is an `InvalidArgumentError` with the reference's message.
"""
from typing import Iterator, Sequence, Tuple, Union

import numpy as np
import torch
from torch.utils.data import IterableDataset

from compression_b200._lib import InvalidArgumentError

__all__ = ["Y4MDataset"]

_FRAME = b"FRAME\n"


def _parse_header(header: bytes, filename: str):
  """-> (width, height, chroma subsampling 2 | 1); `header` ends with its newline (y4m_dataset_kernels.cc:300-400)."""
  body = header[:-1]
  if not body.startswith(b"YUV4MPEG2"):
    raise InvalidArgumentError(f"Input file '{filename}' does not have a YUV4MPEG2 marker.")
  body = body[len(b"YUV4MPEG2"):]
  width = height = 0
  chroma = None
  while body:
    if len(body) < 2 or body[:1] != b" ":
      raise InvalidArgumentError(
          f"Input file '{filename}' has an invalid Y4M header. Remaining header: '{body.decode(errors='replace')}'.")
    key, body = body[1:2], body[2:]
    if key in (b"W", b"H"):
      n = 0
      while n < len(body) and body[n:n + 1].isdigit():
        n += 1
      what = "width" if key == b"W" else "height"
      value = int(body[:n]) if n else 0
      if value <= 0:
        raise InvalidArgumentError(
            f"Input file '{filename}' has an invalid {what} specifier '{body[:n].decode(errors='replace')}'.")
      if key == b"W":
        width = value
      else:
        height = value
      body = body[n:]
    elif key == b"C":
      for tag, sub in ((b"420jpeg", 2), (b"420", 2), (b"444", 1)):
        if body.startswith(tag):
          chroma, body = sub, body[len(tag):]
          break
      else:
        raise InvalidArgumentError(
            f"Input file '{filename}' has an unsupported chroma format '{body.split(b' ')[0].decode(errors='replace')}'.")
    elif key == b"I":
      if not body.startswith(b"p"):
        raise InvalidArgumentError(f"Input file '{filename}' is not in progressive format.")
      body = body[1:]
    else:                                   # frame rate, aspect ratio, comments ...: skipped up to the next field
      at = body.find(b" ")
      body = b"" if at < 0 else body[at:]
  if not width:
    raise InvalidArgumentError(f"Input file '{filename}' has no width specifier.")
  if not height:
    raise InvalidArgumentError(f"Input file '{filename}' has no height specifier.")
  if chroma is None:
    raise InvalidArgumentError(f"Input file '{filename}' has no chroma format specifier.")
  if chroma == 2 and (width & 1 or height & 1):
    raise InvalidArgumentError(f"Input file '{filename}' has 4:2:0 chroma format, but odd width or height.")
  return width, height, chroma


class Y4MDataset(IterableDataset):
  """Frames of one or more '.y4m' files as `(y, cbcr)` uint8 tensor pairs, all files concatenated."""

  def __init__(self, filenames: Union[str, Sequence[str]]):
    super().__init__()
    self.filenames = [filenames] if isinstance(filenames, (str, bytes)) else list(filenames)

  def __iter__(self) -> Iterator[Tuple[torch.Tensor, torch.Tensor]]:
    for filename in self.filenames:
      with open(filename, "rb") as f:
        header = f.readline()
        if not header.endswith(b"\n"):
          raise InvalidArgumentError(f"Input file '{filename}' does not contain a complete Y4M header.")
        width, height, sub = _parse_header(header, str(filename))
        cw, ch = width // sub, height // sub
        n_y, n_c = width * height, cw * ch
        size = len(_FRAME) + n_y + 2 * n_c
        pos = len(header)
        while True:
          buf = f.read(size)
          if not buf:
            break                                                     # clean end of file: on to the next one
          if len(buf) < size:
            raise InvalidArgumentError(
                f"Input file '{filename}' has an incomplete or unsupported frame at byte {pos}. Expected to read "
                f"{size} bytes, only {len(buf)} were available.")
          if not buf.startswith(_FRAME):
            raise InvalidArgumentError(
                f"Input file '{filename}' has a FRAME marker at byte {pos} which is either invalid or has "
                "unsupported frame parameters.")
          planes = np.frombuffer(buf, np.uint8, offset=len(_FRAME))
          y = planes[:n_y].reshape(height, width, 1).copy()
          cbcr = np.stack([planes[n_y:n_y + n_c], planes[n_y + n_c:]], axis=-1).reshape(ch, cw, 2)
          pos += size
          yield torch.from_numpy(y), torch.from_numpy(cbcr)
