"""Host-side logic that needs no GPU: the drop-in module surface (constructor, state_dict ABI, seeded init,
pickling, checkpoint loading), the batch partitioner and the world_size-2 gloo score gather."""
import io
import os
import pickle

import numpy as np
import pytest
import torch

import bbbp_b200
from bbbp_b200 import partition_batches
from conftest import load_golden
from oracle import nets

VARIANT_DIMS = [
    ("tcnn", 167, 128), ("tcnn_nofusion", 167, 128), ("tcnn_big", 167, 128), ("mlp", 64, 128), ("mlp_morgan", 128, 256),
    ("mlp_rdkit", 64, 128), ("mlp_more", 64, 128),
]


@pytest.mark.parametrize("variant,fp_dim,img_side", VARIANT_DIMS)
def test_state_dict_abi_and_seeded_init_match_reference_layout(variant, fp_dim, img_side):
    torch.manual_seed(5)
    ours = bbbp_b200.build(variant, fp_dim, img_side)
    torch.manual_seed(5)
    ref = nets.build(variant, fp_dim, img_side)       # pinned to the reference classes in test_oracle_pinning
    a, b = ours.state_dict(), ref.state_dict()
    assert list(a) == list(b)
    for k in a:
        assert a[k].shape == b[k].shape and a[k].dtype == b[k].dtype and torch.equal(a[k], b[k]), k
    ref.load_state_dict(a, strict=True)
    ours.load_state_dict(b, strict=True)


def test_parameter_counts():
    count = lambda m: sum(p.numel() for p in m.parameters())
    assert count(bbbp_b200.MixedInputModel(167, 128)) == 13_464_087
    assert count(bbbp_b200.MixedInputModelBig(167, 128)) == 46_203_559
    assert count(bbbp_b200.MixedInputModelMLP(64, 128)) == 198_149       # best_nn_model_maccs.pth
    assert count(bbbp_b200.MixedInputModelMLP(128, 256)) == 222_725      # best_nn_model.pth


def test_encoder_head_rule():
    assert bbbp_b200.encoder_heads(167) == 1          # 167 is prime
    assert bbbp_b200.encoder_heads(2048) == 256
    assert bbbp_b200.encoder_heads(167, 8) == 1
    assert bbbp_b200.encoder_heads(2048, 8) == 8
    m = bbbp_b200.MixedInputModel(2048 // 8, 128)     # small stand-in: 256 -> 32 heads
    assert m.fingerprint_transformer.layers[0].self_attn.num_heads == 32


@pytest.mark.parametrize("name,fp_dim,img_dim", [("mlp_ckpt_maccs", 64, 128), ("mlp_ckpt_morgan", 128, 256)])
def test_shipped_checkpoints_load_strict(name, fp_dim, img_dim):
    g = load_golden(name)
    state = {k[len("param:"):]: torch.from_numpy(g[k]) for k in g.files if k.startswith("param:")}
    model = bbbp_b200.MixedInputModelMLP(fp_dim, img_dim)
    model.load_state_dict(state, strict=True)
    buf = io.BytesIO()
    torch.save(model.state_dict(), buf)               # Models/..._opt.py:179
    buf.seek(0)
    again = bbbp_b200.MixedInputModelMLP(fp_dim, img_dim)
    again.load_state_dict(torch.load(buf), strict=True)


def test_whole_model_pickles():
    model = bbbp_b200.MixedInputModelMLP(64, 128)      # 20250113.py:243 pickles the whole module
    clone = pickle.loads(pickle.dumps(model))
    for (k, a), (_, b) in zip(model.state_dict().items(), clone.state_dict().items()):
        assert torch.equal(a, b), k


def test_pickle_and_deepcopy_drop_runtime_caches():
    """The reference pickles the model right AFTER its eval loop (20250113.py:229-244), i.e. once CUDA graphs, streams and
    staging buffers hang off the instance: those run-time caches must not travel (a CUDAGraph cannot be pickled)."""
    import copy
    model = bbbp_b200.MixedInputModel(167, 128)
    unpicklable = lambda: None                                   # stands in for torch.cuda.CUDAGraph / Stream / Event
    model._graphs = {"k": (unpicklable,)}
    model._chunk_graphs = {"k": (unpicklable,)}
    model._host_pipe = ("key", unpicklable)
    model._side_stream = unpicklable
    model._sig_tensors = list(model.parameters())
    model.attention_fusion._stack_cache = ("key", unpicklable)
    for clone in (pickle.loads(pickle.dumps(model)), copy.deepcopy(model)):
        assert clone._graphs is None and clone._chunk_graphs is None and clone._side_stream is None
        assert getattr(clone, "_host_pipe", None) is None and clone._sig_tensors is None
        assert getattr(clone.attention_fusion, "_stack_cache", None) is None
        for (k, a), (_, b) in zip(model.state_dict().items(), clone.state_dict().items()):
            assert torch.equal(a, b), k
    assert model._graphs is not None                             # the original keeps its caches


def test_cpu_tensors_are_refused_loudly():
    model = bbbp_b200.MixedInputModelMLP(64, 128)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model(torch.randn(2, 64), torch.randn(2, 128))


def test_product_never_imports_the_oracle():
    pkg = os.path.dirname(bbbp_b200.__file__)
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "import oracle" not in src and "from oracle" not in src, fn


@pytest.mark.parametrize("n,bs,world", [(0, 256, 4), (1, 256, 8), (1058, 256, 1), (1058, 256, 2), (1058, 256, 8),
                                        (7807, 32, 4), (10_000_000, 256, 8), (255, 256, 2), (512, 256, 3)])
def test_partitioner_covers_every_molecule_once_on_batch_boundaries(n, bs, world):
    spans = [partition_batches(n, bs, world, r) for r in range(world)]
    assert spans[0][0] == 0 and spans[-1][1] == n
    for (a0, a1), (b0, b1) in zip(spans, spans[1:]):
        assert a1 == b0 and a0 <= a1
    for a, b in spans:
        assert a % bs == 0                      # every rank starts on a reference-batch boundary
        assert b % bs == 0 or b == n            # only the global tail batch is ragged
    nb = -(-n // bs)
    sizes = [-(-(b - a) // bs) for a, b in spans]
    assert sum(sizes) == nb and max(sizes) - min(sizes) <= 1


def test_partitioner_rejects_bad_arguments():
    with pytest.raises(ValueError):
        partition_batches(10, 0, 1, 0)
    with pytest.raises(ValueError):
        partition_batches(10, 4, 2, 2)


def _gather_worker(rank, world, port, n, bs, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        a, b = partition_batches(n, bs, world, rank)
        local = torch.arange(a, b, dtype=torch.float32) * 0.5      # "score" of molecule i is i/2
        full = bbbp_b200.gather_scores(local, n, bs)
        np.save(os.path.join(out_dir, f"r{rank}.npy"), full.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n,bs", [(1058, 256), (600, 32), (3, 8)])
def test_score_gather_world2_gloo(tmp_path, n, bs):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_gather_worker, args=(2, port, n, bs, str(tmp_path)), nprocs=2, join=True)
    want = np.arange(n, dtype=np.float32) * 0.5
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"r{r}.npy"), want)


def _avg_grad_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)
        params = [torch.nn.Parameter(torch.zeros(5, 3)), torch.nn.Parameter(torch.zeros(7)), torch.nn.Parameter(torch.zeros(2))]
        params[0].grad = torch.full((5, 3), float(rank + 1))
        params[1].grad = torch.arange(7, dtype=torch.float32) * (rank + 1)
        bbbp_b200.average_gradients(params)          # params[2] has no gradient: skipped
        np.save(os.path.join(out_dir, f"g{rank}.npy"), torch.cat([params[0].grad.reshape(-1), params[1].grad]).numpy())
    finally:
        dist.destroy_process_group()


def test_gradient_averaging_world2_gloo(tmp_path):
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_avg_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    want = np.concatenate([np.full(15, 1.5, dtype=np.float32), np.arange(7, dtype=np.float32) * 1.5])
    for r in range(2):
        np.testing.assert_array_equal(np.load(tmp_path / f"g{r}.npy"), want)


def _flat_grad_worker(rank, world, port, out_dir):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.manual_seed(0)                                    # identical replicas
        net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 1))
        unused = torch.nn.Parameter(torch.ones(4))              # never reached by backward: its bucket is reduced as zeros
        flat = bbbp_b200.FlatGradients(list(net.parameters()) + [unused], bucket_bytes=32)      # several tiny buckets
        assert len(flat.buckets) >= 3 and all(p.grad.data_ptr() >= flat.flat.data_ptr() for p in net.parameters())
        out = []
        for step in range(2):
            g = torch.Generator().manual_seed(10 * step + rank)         # a different micro-batch per rank
            x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
            flat.zero()
            torch.nn.functional.mse_loss(net(x), y).backward()          # hooks launch the bucket all-reduces during backward
            flat.synchronize()
            out.append(flat.flat.clone())
        np.save(os.path.join(out_dir, f"f{rank}.npy"), torch.stack(out).numpy())
        flat.close()
    finally:
        dist.destroy_process_group()


def test_flat_gradient_buckets_average_during_backward_world2_gloo(tmp_path):
    """SURVEY 8e data-parallel training: every p.grad is a view into ONE persistent buffer, buckets are all-reduced as
    backward completes them; the result equals the mean of the two replicas' micro-batch gradients."""
    import socket
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_flat_grad_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    got = [np.load(tmp_path / f"f{r}.npy") for r in range(2)]
    np.testing.assert_array_equal(got[0], got[1])
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3), torch.nn.Linear(3, 1))
    for step in range(2):
        grads = []
        for rank in range(2):
            g = torch.Generator().manual_seed(10 * step + rank)
            x, y = torch.randn(8, 6, generator=g), torch.randn(8, 1, generator=g)
            net.zero_grad()
            torch.nn.functional.mse_loss(net(x), y).backward()
            grads.append(torch.cat([p.grad.reshape(-1) for p in reversed(list(net.parameters()))]))
        want = torch.cat([torch.zeros(4), (grads[0] + grads[1]) / 2])       # buffer layout: reverse parameter order
        np.testing.assert_allclose(got[0][step], want.numpy(), rtol=1e-6, atol=1e-7)


def test_device_batch_feeder_reproduces_dataloader_order():
    """same batches, same order as the reference's MixedDataset + DataLoader(shuffle=True) under the same
    torch seed, for several epochs (the DataLoader draws one seed per epoch from the global generator)."""
    from torch.utils.data import DataLoader, Dataset
    n, bs = 77, 32
    fps = np.arange(n * 5, dtype=np.float64).reshape(n, 5)
    imgs = np.arange(n * 12, dtype=np.float32).reshape(n, 12) * 0.5
    ys = np.arange(n, dtype=np.float64) * 0.25

    class MixedDataset(Dataset):          # the reference's dataset shape (20250113.py:31-45), re-stated for the test
        def __len__(self):
            return n

        def __getitem__(self, i):
            return (torch.tensor(fps[i], dtype=torch.float32), torch.tensor(imgs[i], dtype=torch.float32),
                    torch.tensor(ys[i], dtype=torch.float32))

    torch.manual_seed(42)
    want = [[tuple(t.clone() for t in batch) for batch in DataLoader(MixedDataset(), batch_size=bs, shuffle=True)] for _ in range(3)]
    torch.manual_seed(42)
    feeder = bbbp_b200.DeviceBatchFeeder(fps, imgs, ys, batch_size=bs, shuffle=True, device="cpu")
    assert len(feeder) == 3
    for epoch in range(3):
        got = list(feeder)
        assert len(got) == len(want[epoch])
        for (a, b, c), (x, y, z) in zip(got, want[epoch]):
            assert torch.equal(a, x) and torch.equal(b, y) and torch.equal(c, z)
    plain = list(bbbp_b200.DeviceBatchFeeder(fps, imgs, ys, batch_size=bs, shuffle=False, device="cpu"))
    assert torch.equal(plain[-1][2], torch.tensor(ys[64:], dtype=torch.float32))


def test_stack_columns_matches_reference_vstack_transpose():
    """20250113.py:403: ``np.vstack([nn, rf, xgb, cat]).T``."""
    import numpy as np
    import torch
    import bbbp_b200
    rng = np.random.default_rng(1)
    nn_col, rf, xgb = rng.normal(size=7).astype(np.float32), rng.normal(size=7), rng.normal(size=7)
    want = np.vstack([nn_col, rf, xgb]).T
    np.testing.assert_array_equal(bbbp_b200.stack_columns(nn_col, rf, xgb), want)
    np.testing.assert_array_equal(bbbp_b200.stack_columns(torch.from_numpy(nn_col), rf, xgb), want)
    import pytest
    with pytest.raises(ValueError):
        bbbp_b200.stack_columns(nn_col, rf[:5])


def test_host_pipeline_chunk_spans_cover_whole_batches():
    """predict_from_host's chunk schedule: contiguous, whole reference batches per chunk, the ragged tail on the last span."""
    import bbbp_b200
    spans = bbbp_b200.TransformerCnnModel._pipeline_spans
    for n, chunk, bs in [(8192, 1024, 256), (1058, 64, 32), (1058, 96, 32), (300, 2048, 256), (31, 64, 32), (0, 64, 32), (32, 64, 32),
                         (10_000, 1024, 256)]:
        sp = spans(n, chunk, bs)
        if n == 0:
            assert sp == []
            continue
        assert sp[0][0] == 0 and sp[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(sp, sp[1:]))
        assert all(a % bs == 0 for a, _ in sp)                      # every chunk starts on a reference-batch boundary
        assert all(b - a <= max(chunk, bs) for a, b in sp)          # and fits its staging slot
        assert all((b - a) % bs == 0 for a, b in sp[:-1])           # only the last span may carry the ragged tail batch


def test_host_pipeline_chunk_schedules():
    """A schedule of chunk lengths (short first chunk, the last length repeating) keeps the same invariants."""
    import bbbp_b200
    spans = bbbp_b200.TransformerCnnModel._pipeline_spans
    assert spans(16384, (2048, 6144, 8192), 256) == [(0, 2048), (2048, 8192), (8192, 16384)]
    assert spans(16384 + 100, (2048, 14336), 256) == [(0, 2048), (2048, 16384), (16384, 16484)]
    assert spans(5000, (1024, 2048), 256) == [(0, 1024), (1024, 3072), (3072, 5000)]      # tail batch rides on the last span
    assert spans(100, (1024, 2048), 256) == [(0, 100)]
    for n, sched, bs in [(10_000, (512, 1024, 4096), 256), (1058, (64, 96), 32), (777, (256,), 256)]:
        sp = spans(n, sched, bs)
        assert sp[0][0] == 0 and sp[-1][1] == n and all(a[1] == b[0] for a, b in zip(sp, sp[1:]))
        assert all(a % bs == 0 for a, _ in sp) and all(b - a <= max(max(sched), bs) for a, b in sp)


def test_graphed_train_step_rejects_foreign_optimizers():
    """The captured update reads its scalars from device memory: only bbbp_b200.AdamW provides that entry point."""
    import pytest
    import torch
    import bbbp_b200
    model = bbbp_b200.MixedInputModelMLP(64, 128)
    with pytest.raises(TypeError, match="bbbp_b200.AdamW"):
        bbbp_b200.GraphedTrainStep(model, torch.optim.AdamW(model.parameters()), bbbp_b200.MSELoss())


def test_pack_fingerprint_bits_is_little_endian_packbits():
    import numpy as np
    import bbbp_b200
    from oracle import preprocess
    rng = np.random.default_rng(3)
    bits = (rng.random((5, 167)) < 0.3).astype(np.int64)
    packed = bbbp_b200.pack_fingerprint_bits(bits).numpy()
    np.testing.assert_array_equal(packed, preprocess.pack_bits(bits))
    np.testing.assert_array_equal(preprocess.unpack_bits(packed, 167), bits)
