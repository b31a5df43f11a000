"""Training driver host logic on the CPU with a stand-in network: options, optimiser / scheduler table,
gradient accumulation, the validation means, scalar files, and the 2-rank (gloo) epoch against the single-process one."""
import json
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from geobi_gnn_b200.data import Data


class _Stub(torch.nn.Module):
    """Same call shape as DualGNN: [Data_v, Data_f] -> (vert_p [Nv,3], norm_p [Nf,3] unit, None)."""

    def __init__(self):
        super().__init__()
        self.lin_v = torch.nn.Linear(6, 3)
        self.lin_f = torch.nn.Linear(6, 3)

    def forward(self, data):
        return self.lin_v(data[0].x), torch.nn.functional.normalize(self.lin_f(data[1].x), dim=1), None


def _samples(n, seed=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for i in range(n):
        nv, nf = 5 + i, 9 + 2 * i
        out.append((Data(x=torch.randn(nv, 6, generator=g), y=torch.randn(nv, 3, generator=g)),
                    Data(x=torch.randn(nf, 6, generator=g), y=torch.nn.functional.normalize(torch.randn(nf, 3, generator=g), dim=1))))
    return out


def _opt(**kw):
    from geobi_gnn_b200 import train
    opt = train.parse_arguments(["--data_type=Synthetic", "--flag=t", "--gpu=-1"])
    opt.__dict__.update(kw)
    return opt


def test_options_match_train_dual():
    from geobi_gnn_b200 import train
    opt = train.parse_arguments(["--data_type=Kinect_v2", "--flag=x", "--gpu=3", "--lr_step", "5", "9", "--extra=[1,2]", "--name=abc"])
    assert (opt.sub_size, opt.filter_patch_count, opt.wei_param, opt.batch_size, opt.lr, opt.lr_sch) == (20000, 100, 2, 1, 1e-3, "lmd")
    assert opt.lr_step == [5, 9] and opt.extra == [1, 2] and opt.name == "abc"
    assert opt.force_depth is True and opt.pool_type == "max" and opt.loss_v == opt.loss_n == "L1"
    assert train.parse_arguments(["--data_type=Synthetic", "--flag=x", "--gpu=0"]).force_depth is False
    with pytest.raises(SystemExit):
        train.parse_arguments(["--flag=x"])


def test_optimizer_and_scheduler_table():
    from geobi_gnn_b200 import train
    net = _Stub()
    o = train.make_optimizer(_opt(optimizer="adam", beta1=0.8, beta2=0.9, weight_decay=0.1), net.parameters())
    assert isinstance(o, torch.optim.Adam) and o.defaults["betas"] == (0.8, 0.9) and o.defaults["weight_decay"] == 0.1
    assert isinstance(train.make_optimizer(_opt(optimizer="sgd"), net.parameters()), torch.optim.SGD)
    r = train.make_optimizer(_opt(optimizer="rmsprop"), net.parameters())
    assert isinstance(r, torch.optim.RMSprop) and r.defaults["alpha"] == 0.9
    with pytest.raises(ValueError):
        train.make_optimizer(_opt(optimizer="lion"), net.parameters())
    sch = torch.optim.lr_scheduler
    for name, cls in (("step", sch.StepLR), ("multi_step", sch.MultiStepLR), ("exp", sch.ExponentialLR), ("auto", sch.ReduceLROnPlateau),
                      ("lmd", sch.LambdaLR), ("anything", sch.LambdaLR)):
        opt = _opt(lr_sch=name, lr_decay=0.5, lr_step=[4, 8], lr=0.1)
        o = train.make_optimizer(opt, net.parameters())
        assert isinstance(train.make_scheduler(opt, o), cls)
    opt = _opt(lr_sch="lmd", lr_decay=0.5, lr_step=[4], lr=0.1)
    o = train.make_optimizer(opt, net.parameters())
    s = train.make_scheduler(opt, o)
    for _ in range(6):
        o.step()
        s.step()
    assert abs(o.param_groups[0]["lr"] - 0.1 * 0.5 ** (6 / 4)) < 1e-12          # train_dual.py:178-180


def test_accumulation_equals_one_averaged_step(tmp_path):
    """train_dual.py:211-218: `batch_size` backward passes of loss / batch_size, then one optimiser step; the tail of an epoch
    that does not fill a batch still steps.  Scalars land in scalars.jsonl at upstream's iteration numbers."""
    from geobi_gnn_b200 import network, train
    samples = _samples(5)
    opt = _opt(batch_size=2, optimizer="sgd", momentum=0.0, lr=0.1, loss_v="L2", loss_v_scale=2.0, loss_n_scale=0.5)
    torch.manual_seed(1)
    net = _Stub()
    ref = _Stub()
    ref.load_state_dict(net.state_dict())
    writer = train.ScalarWriter(str(tmp_path / "train"), tensorboard=False)
    last = train.train_epoch(net, train.make_optimizer(opt, net.parameters()), samples, opt, "cpu", writer, first_iteration=100)
    writer.close()
    o = train.make_optimizer(opt, ref.parameters())
    for group in ([0, 1], [2, 3], [4]):
        o.zero_grad()
        total = 0
        for i in group:
            vp, np_, _ = ref(samples[i])
            total = total + network.dual_loss(network.loss_v(vp, samples[i][0].y, "L2"), network.loss_n(np_, samples[i][1].y, "L1"), 2.0, 0.5) / 2
        total.backward()
        o.step()
    for a, b in zip(net.parameters(), ref.parameters()):
        assert torch.allclose(a, b, rtol=1e-6, atol=1e-7)
    rows = [json.loads(l) for l in open(tmp_path / "train" / "scalars.jsonl")]
    assert [r["step"] for r in rows if r["tag"] == "dual_loss"] == [101, 103, 104]
    assert {r["tag"] for r in rows} == {"loss_v", "loss_f", "dual_loss", "error_v", "error_f"}
    assert len(last) == 5 and abs(last[2] - (2.0 * last[0] + 0.5 * last[1])) < 1e-5


def test_validation_means_are_node_weighted():
    from geobi_gnn_b200 import network, train
    samples = _samples(4, seed=3)
    net = _Stub()
    got = train.evaluate(net, samples, _opt(), "cpu")
    assert not net.training
    with torch.no_grad():
        lv = lf = ev = ef = cv = cf = 0
        for d in samples:
            vp, np_, _ = net(d)
            lv += network.loss_v(vp, d[0].y, "L1") * d[0].y.shape[0]
            lf += network.loss_n(np_, d[1].y, "L1") * d[1].y.shape[0]
            ev += network.error_v(vp, d[0].y) * d[0].y.shape[0]
            ef += network.error_n(np_, d[1].y) * d[1].y.shape[0]
            cv += d[0].y.shape[0]
            cf += d[1].y.shape[0]
    want = (float(lv / cv), float(lf / cf), float(ev / cv), float(ef / cf))
    assert all(abs(g - w) <= 1e-5 * abs(w) for g, w in zip(got, want))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geobi_gnn_b200 import train
    samples = _samples(8, seed=5)
    opt = _opt(batch_size=2, optimizer="adam", lr=0.01)
    torch.manual_seed(2)
    net = _Stub()
    train.train_epoch(net, train.make_optimizer(opt, net.parameters()), samples[rank::world], opt, "cpu")
    if rank == 0:
        q.put([p.detach().tolist() for p in net.parameters()])      # plain lists: no shared-memory handles
    dist.destroy_process_group()


def test_two_rank_epoch_equals_single_process_with_doubled_batch():
    """World 2 x batch_size 2 over samples dealt round-robin == one process, batch_size 4, same order."""
    from geobi_gnn_b200 import train
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    samples = _samples(8, seed=5)
    opt = _opt(batch_size=4, optimizer="adam", lr=0.01)
    torch.manual_seed(2)
    net = _Stub()
    train.train_epoch(net, train.make_optimizer(opt, net.parameters()), samples, opt, "cpu")
    for a, b in zip(got, net.parameters()):
        assert torch.allclose(torch.tensor(a), b, rtol=1e-5, atol=1e-6)


def test_whole_training_run_on_a_cached_data_set(tmp_path):
    """train.train end to end on the CPU: a cached data set (nothing to build), a stand-in network, 3 epochs with accumulation
    over 2 and a step schedule - run directory, params / model files, log, scalar files, best-model rule, restored stdout."""
    import sys
    import numpy as np
    from geobi_gnn_b200 import checkpoint, dataset, meshio, synth, train
    from oracle import ref_dataset
    from tests import util
    root, logs = tmp_path / "dataset", tmp_path / "log"
    for split, names in (("train", ("a", "b", "c")), ("test", ("t",))):
        base = root / "Synthetic" / split
        for sub in ("original", "noisy", "processed_data"):
            (base / sub).mkdir(parents=True)
        for i, name in enumerate(names):
            mesh_n, mesh_o = util.noisy_icosphere(2, seed=i)
            dd = ref_dataset.process_one_submesh(mesh_n, name, mesh_o)
            ref_dataset.attach_normalisation(dd, mesh_n.points, mesh_n.ev)
            meshio.write_obj(base / "original" / f"{name}.obj", mesh_o.points, mesh_o.fv)
            meshio.write_obj(base / "noisy" / f"{name}_n1.obj", mesh_n.points, mesh_n.fv)
            dataset.save_dual_data(tuple(util.data_to(d, "cpu") for d in dd), base / "processed_data" / f"{name}_n1.pt")
    opt = train.parse_arguments(["--data_type=Synthetic", "--flag=cpu", "--gpu=-1", "--seed=3", "--max_epoch=3", "--batch_size=2",
                                 "--lr_sch=step", "--lr_decay=0.5", "--lr_step", "1", "--lr=0.05", "--optimizer=sgd"])
    stdout = sys.stdout
    params_file = train.train(opt, dataset_root=str(root), log_root=str(logs), tensorboard=False, net_factory=lambda o: _Stub())
    assert sys.stdout is stdout
    run_dir = os.path.dirname(params_file)
    assert os.path.dirname(os.path.dirname(run_dir)) == str(logs) and os.path.basename(os.path.dirname(run_dir)) == "GeoBi-GNN_Synthetic_cpu"
    saved = checkpoint.load_params(params_file)
    assert saved.flag.startswith("GeoBi-GNN_Synthetic_cpu_") and saved.model_name == "GeoBi-GNN_Synthetic_model.pth" and saved.seed == 3
    state = torch.load(os.path.join(run_dir, saved.model_name), weights_only=True)
    assert set(state) == {"lin_v.weight", "lin_v.bias", "lin_f.weight", "lin_f.bias"}
    log = open(os.path.join(run_dir, "training_info.txt")).read()
    assert "Training set:    3 samples" in log and "Testing set:     1 samples" in log and "Epoch   0" in log and "best error" in log
    rows = [json.loads(l) for l in open(os.path.join(run_dir, "train", "scalars.jsonl"))]
    steps = [r["step"] for r in rows if r.get("tag") == "dual_loss"]
    assert steps == [-3 + 1, -3 + 2, 1, 2, 3 + 1, 3 + 2]            # len(loader) * (epoch - 1) + step, as upstream counts (train_dual.py:200)
    evals = [r for r in (json.loads(l) for l in open(os.path.join(run_dir, "test", "scalars.jsonl"))) if r.get("tag") == "error_f"]
    assert len(evals) == 3 and all(np.isfinite(r["value"]) for r in evals)
    lrs = [float(x.split("lr:")[1].split()[0]) for x in log.splitlines() if x.startswith("Epoch")]
    assert lrs and abs(lrs[0] - 0.05) < 1e-12                                           # epoch 0 runs at the initial rate
    with pytest.raises(RuntimeError, match="CUDA"):
        if torch.cuda.is_available():
            raise RuntimeError("CUDA present: the refusal below is only for machines without one")
        train.train(train.parse_arguments(["--data_type=Synthetic", "--flag=x", "--gpu=-1"]), dataset_root=str(root), log_root=str(logs))


def _make_cached_sets(root, n_train=4):
    from geobi_gnn_b200 import dataset, meshio
    from oracle import ref_dataset
    from tests import util
    for split, names in (("train", [f"m{i}" for i in range(n_train)]), ("test", ["t"])):
        base = os.path.join(root, "Synthetic", split)
        for sub in ("original", "noisy", "processed_data"):
            os.makedirs(os.path.join(base, sub), exist_ok=True)
        for i, name in enumerate(names):
            mesh_n, mesh_o = util.noisy_icosphere(2, seed=i)
            dd = ref_dataset.process_one_submesh(mesh_n, name, mesh_o)
            ref_dataset.attach_normalisation(dd, mesh_n.points, mesh_n.ev)
            meshio.write_obj(os.path.join(base, "original", f"{name}.obj"), mesh_o.points, mesh_o.fv)
            meshio.write_obj(os.path.join(base, "noisy", f"{name}_n1.obj"), mesh_n.points, mesh_n.fv)
            dataset.save_dual_data(tuple(util.data_to(d, "cpu") for d in dd), os.path.join(base, "processed_data", f"{name}_n1.pt"))


def _train_worker(rank, world, port, root, logs, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from geobi_gnn_b200 import train
    opt = train.parse_arguments(["--data_type=Synthetic", "--flag=ddp", "--gpu=-1", "--max_epoch=2", "--batch_size=1", "--lr=0.05",
                                 "--optimizer=sgd"])                       # no --seed: rank 0 draws it, everyone must use it
    made = []
    params_file = train.train(opt, dataset_root=root, log_root=logs, tensorboard=False, net_factory=lambda o: made.append(_Stub()) or made[-1])
    q.put((rank, params_file, opt.seed, opt.flag, [p.detach().tolist() for p in made[0].parameters()]))
    dist.destroy_process_group()


def test_two_rank_training_run_shares_seed_run_directory_and_weights(tmp_path):
    """train.train under a 2-rank process group (gloo): rank 0's seed and time stamp are everyone's, one run directory is
    written (by rank 0), both ranks start from the same weights and stay in step (identical parameters at the end), and the saved
    model is loadable.  (Equality with a single process at the doubled batch is checked at train_epoch level above; whole runs
    differ in the augmentation draws, which every rank takes from its own copy of the seeded numpy stream.)"""
    from geobi_gnn_b200 import checkpoint, train
    root, logs = str(tmp_path / "dataset"), str(tmp_path / "log")
    _make_cached_sets(root, n_train=4)
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_train_worker, args=(r, world, port, root, logs, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted(q.get(timeout=240) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, params0, seed0, flag0, w0), (_, params1, seed1, flag1, w1) = got
    assert params0 == params1 and seed0 == seed1 and flag0 == flag1
    runs = os.listdir(os.path.join(logs, "GeoBi-GNN_Synthetic_ddp"))
    assert len(runs) == 1
    for a, b in zip(w0, w1):
        assert torch.allclose(torch.tensor(a), torch.tensor(b), rtol=1e-6, atol=1e-7)
    torch.manual_seed(seed0)
    untouched = _Stub()
    assert any(not torch.allclose(torch.tensor(a), p) for a, p in zip(w0, untouched.parameters()))      # it did train
    saved = checkpoint.load_params(params0)
    state = torch.load(os.path.join(os.path.dirname(params0), saved.model_name), weights_only=True)
    assert set(state) == {"lin_v.weight", "lin_v.bias", "lin_f.weight", "lin_f.bias"}
    rows = [json.loads(l) for l in open(os.path.join(os.path.dirname(params0), "train", "scalars.jsonl"))]
    assert len([r for r in rows if r.get("tag") == "dual_loss"]) == 2 * 2          # 2 epochs x (4 samples / 2 ranks) optimiser steps
