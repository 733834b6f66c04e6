"""Host-side reservoir of (left, right, gt_disp) samples for the experience-replay term of the adaptation step
(BASELINE.json configs[4]: "reservoir-replay batches").  Same policy as the reference's StereoReservoir
(adaptive_stereo/utils/stereo_reservoir.py:5-64: Algorithm R, duplicates by index rejected), holding device tensors; the draw
of a replay sample is seeded so that every run (and every rank, with its own seed) is reproducible."""
import random


class ReplayReservoir:
  def __init__(self, max_size, seed=0):
    self.max_size = max_size
    self.buf = []                        # [index, left, right, gt]
    self.indices = set()
    self.i = 0                           # items streamed so far
    self.rng = random.Random(seed)

  def add(self, left, right, gt, index):
    self.i += 1
    if index in self.indices:
      return False
    item = [index, left.detach(), right.detach(), gt.detach()]
    if len(self.buf) < self.max_size:
      self.buf.append(item)
      self.indices.add(index)
      return True
    j = self.rng.randint(1, self.i)      # replace with gradually decreasing probability (Algorithm R)
    if j <= self.max_size:
      self.indices.discard(self.buf[j - 1][0])
      self.buf[j - 1] = item
      self.indices.add(index)
      return True
    return False

  def size(self):
    return len(self.buf)

  def sample(self):
    """One stored (left, right, gt) triple, uniformly."""
    _, l, r, g = self.buf[self.rng.randrange(len(self.buf))]
    return l, r, g
