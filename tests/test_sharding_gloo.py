"""CPU, world_size 2 over gloo: the N>1 path of the hot path = batch sharding + one table broadcast."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import util


def _free_port():
  with socket.socket() as s:
    s.bind(("127.0.0.1", 0))
    return s.getsockname()[1]


def _tables():
  rng = np.random.default_rng(3)
  cdfs = [util.random_cdf(rng, 10 + c, 12) for c in range(5)]
  lookup = util.make_lookup_1d(cdfs, [12] * 5, [True] * 5)
  return torch.from_numpy(lookup), torch.arange(-3, 2, dtype=torch.int32), torch.linspace(-.4, .4, 5)


def _worker(rank, world, port, q):
  os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
  dist.init_process_group("gloo", rank=rank, world_size=world)
  try:
    from compression_b200 import entropy_models as em
    from compression_b200 import sharding
    if rank == 0:
      cdf, coff, qoff = _tables()
      model = em.ContinuousBatchedEntropyModel(prior_shape=(5,), coding_rank=3, compression=True, cdf=cdf,
                                               cdf_offset=coff, quantization_offset=qoff)
    else:
      model = em.ContinuousBatchedEntropyModel(prior_shape=(5,), coding_rank=3, compression=True,
                                               cdf_shapes=(1, 1), quantization_offset=True)
    sharding.broadcast_tables(model, src=0)
    lo, hi = sharding.shard_range(257, rank, world)
    # data-parallel training: the per-rank parameter gradients of a layer are averaged with one all-reduce
    lin = torch.nn.Linear(3, 2)
    with torch.no_grad():
      lin.weight.fill_(1.0)
      lin.bias.fill_(0.0)
    lin(torch.full((4, 3), float(rank + 1))).sum().backward()
    n_red = sharding.allreduce_gradients(lin)
    # the CPU half of bench.py's per-rank parity check: every rank codes ITS OWN first cfg2 batch with the oracle
    # and the verdicts are MIN-reduced (round 2: rank 1's batch held an infinite latent and never came back)
    import bench
    import oracle
    fx = bench.load_fixture()
    _, ys = bench.synth_latents(rank, 1)
    value = bench.symbols_of(fx["cdf_offset"], None if fx["qoff"] is None else torch.from_numpy(fx["qoff"]), ys[0])
    O = oracle.best()
    want = O.encode(fx["lookup"], value, None, 2)
    back, ok = O.decode(fx["lookup"], want, value.shape[1], None, 2)
    flag = torch.tensor([1 if (np.array_equal(back, value) and ok.all()) else 0])
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    q.put((rank, model.cdf.clone(), model.cdf_offset.clone(), model.quantization_offset.clone(), lo, hi,
           n_red, lin.weight.grad.clone(), lin.bias.grad.clone(), int(flag.item()), int(np.abs(value).max())))
  finally:
    dist.barrier()
    dist.destroy_process_group()


def _run_world(world):
  port = _free_port()
  ctx = mp.get_context("spawn")
  q = ctx.Queue()
  procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
  for p in procs:
    p.start()
  try:
    got = sorted([q.get(timeout=180) for _ in range(world)], key=lambda t: t[0])
    for p in procs:
      p.join(timeout=120)
      assert p.exitcode == 0
    return got
  finally:
    for p in procs:
      if p.is_alive():
        p.kill()


def test_table_broadcast_and_batch_shards():
  world = 2
  got, last = None, None
  for _ in range(3):  # the rendezvous port is picked by bind-and-release: retry if somebody else grabbed it
    try:
      got = _run_world(world)
      break
    except Exception as e:  # pylint:disable=broad-except
      last = e
  assert got is not None, f"gloo rendezvous failed three times: {last!r}"
  cdf, coff, qoff = _tables()
  for rank, c, o, qo, lo, hi, n_red, gw, gb, parity, sym_max in got:
    assert parity == 1 and sym_max < 1000
    assert torch.equal(c, cdf.to(torch.int32)) and torch.equal(o, coff) and torch.allclose(qo, qoff)
    # rank r's own gradient is 4 (r + 1) per weight and 4 per bias: the average over two ranks is 6 and 4
    assert n_red == 8 and torch.allclose(gw, torch.full((2, 3), 6.0)) and torch.allclose(gb, torch.full((2,), 4.0))
  assert got[0][4] == 0 and got[0][5] == got[1][4] and got[1][5] == 257
  assert abs((got[0][5] - got[0][4]) - (got[1][5] - got[1][4])) <= 1


def test_shard_range_partitions_everything():
  from compression_b200 import sharding
  for n in (0, 1, 7, 256, 2048):
    for world in (1, 2, 3, 8):
      pieces = [sharding.shard_range(n, r, world) for r in range(world)]
      assert pieces[0][0] == 0 and pieces[-1][1] == n
      assert all(a[1] == b[0] for a, b in zip(pieces, pieces[1:]))
      sizes = [hi - lo for lo, hi in pieces]
      assert max(sizes) - min(sizes) <= 1
