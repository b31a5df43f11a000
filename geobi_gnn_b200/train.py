"""Training step with the reference's semantics (train_dual.py:199-218) on N replicas.

The reference accumulates gradients over `batch_size` meshes (one mesh per forward) and then steps Adam.  Here each rank
runs ONE micro-step on its own mesh / patch batch, gradients are summed over ranks with a single flat fp32 all-reduce
(NCCL over NVLink; parallel.allreduce_gradients) and every rank applies the same optimiser step — i.e. `world_size`
plays the role of the reference's `batch_size`, loss pre-scaled by 1/world_size exactly as `train_loss /= opt.batch_size`.
"""
from __future__ import annotations

import torch

from . import network, parallel


def train_step(net, optimizer, dual_data, loss_v="L1", loss_n="L1", v_scale=1.0, n_scale=1.0, world_size=1):
    """One forward / loss / backward / all-reduce / optimiser step.  Returns (loss, error_v, error_n) tensors of this rank."""
    data_v, data_f = dual_data
    y_v, y_f = data_v.y, data_f.y
    optimizer.zero_grad(set_to_none=True)
    vert_p, norm_p, _ = net([data_v, data_f])
    l_v = network.loss_v(vert_p, y_v, loss_v)
    l_f = network.loss_n(norm_p, y_f, loss_n)
    loss = network.dual_loss(l_v, l_f, v_scale=v_scale, n_scale=n_scale)
    (loss / world_size).backward()
    parallel.allreduce_gradients(net.parameters(), average=False)
    optimizer.step()
    with torch.no_grad():
        return loss.detach(), network.error_v(vert_p, y_v), network.error_n(norm_p, y_f)


# ---------------------------------------------------------------------------------------------------------------------
# The driver around the step: train_dual.py's options, optimiser / scheduler table, epoch loop with
# gradient accumulation, validation pass, scalars and run files.  One process per GPU: with a process group the training
# samples are dealt round-robin to ranks and the gradients summed with one all-reduce per optimiser step.
# ---------------------------------------------------------------------------------------------------------------------
import argparse
import ast
import os
import random
import sys
from datetime import datetime

import numpy as np


def parse_arguments(argv=None):
    """train_dual.py:42-96: same options, defaults and derived fields (`force_depth`, `pool_type`).  Unknown `--key=value`
    pairs extend the namespace as upstream's do, but through ast.literal_eval instead of eval."""
    p = argparse.ArgumentParser()
    p.add_argument("--data_type", type=str, required=True, help="Data type for training")
    p.add_argument("--flag", type=str, required=True, help="Training flag")
    p.add_argument("--gpu", type=int, required=True, help="GPU to use")
    p.add_argument("--seed", type=int, default=None)
    p.add_argument("--filter_patch_count", type=int, default=100)
    p.add_argument("--sub_size", type=int, default=20000)
    p.add_argument("--loss_v", type=str, default="L1")
    p.add_argument("--loss_n", type=str, default="L1")
    p.add_argument("--loss_v_scale", type=float, default=1)
    p.add_argument("--loss_n_scale", type=float, default=1)
    p.add_argument("--wei_param", type=int, default=2)
    p.add_argument("--max_epoch", type=int, default=1000)
    p.add_argument("--batch_size", type=int, default=1)
    p.add_argument("--lr_sch", type=str, default="lmd")
    p.add_argument("--lr", type=float, default=0.001)
    p.add_argument("--lr_step", type=int, nargs="+", default=[10])
    p.add_argument("--lr_decay", type=float, default=1)
    p.add_argument("--optimizer", type=str, default="adam")
    p.add_argument("--momentum", type=float, default=0.9)
    p.add_argument("--beta1", type=float, default=0.9)
    p.add_argument("--beta2", type=float, default=0.999)
    p.add_argument("--weight_decay", type=float, default=0)
    p.add_argument("--restore", action="store_true")
    p.add_argument("--model_path", type=str, default=None)
    opt, extra = p.parse_known_args(argv)
    for arg in extra:
        key, _, value = arg[2:].partition("=")
        try:
            opt.__dict__[key] = ast.literal_eval(value)
        except (ValueError, SyntaxError):
            opt.__dict__[key] = value
    opt.force_depth = opt.data_type in ("Kinect_v1", "Kinect_v2")
    opt.pool_type = "max"
    return opt


def make_optimizer(opt, params):
    """train_dual.py:162-167."""
    if opt.optimizer == "sgd":
        return torch.optim.SGD(params, lr=opt.lr, momentum=opt.momentum, weight_decay=opt.weight_decay)
    if opt.optimizer == "rmsprop":
        return torch.optim.RMSprop(params, lr=opt.lr, alpha=0.9)
    if opt.optimizer == "adam":
        return torch.optim.Adam(params, lr=opt.lr, betas=(opt.beta1, opt.beta2), weight_decay=opt.weight_decay)
    raise ValueError(f"optimizer {opt.optimizer!r}: expected sgd, rmsprop or adam")     # upstream would fail later with a NameError


def make_scheduler(opt, optimizer):
    """train_dual.py:169-180; any other name is the `lmd` schedule lr * decay ** (epoch / lr_step[0])."""
    sch = torch.optim.lr_scheduler
    if opt.lr_sch == "step":
        return sch.StepLR(optimizer, step_size=opt.lr_step[0], gamma=opt.lr_decay)
    if opt.lr_sch == "multi_step":
        return sch.MultiStepLR(optimizer, milestones=opt.lr_step, gamma=opt.lr_decay)
    if opt.lr_sch == "exp":
        return sch.ExponentialLR(optimizer, gamma=opt.lr_decay)
    if opt.lr_sch == "auto":
        return sch.ReduceLROnPlateau(optimizer, factor=opt.lr_decay, patience=opt.lr_step[0])
    return sch.LambdaLR(optimizer, lr_lambda=lambda step: opt.lr_decay ** (step / opt.lr_step[0]))


class PrintLogger:
    """train_dual.py:21-34: tee of stdout into the run's training_info.txt."""

    def __init__(self, filename="Default.log"):
        self.terminal = sys.stdout
        self.log = open(filename, "a")

    def write(self, message):
        self.terminal.write(message)
        self.log.write(message)

    def flush(self):
        self.log.flush()

    def close(self):
        self.log.close()
        return self.terminal


class ScalarWriter:
    """The add_scalar / add_text / close part of tensorboard's SummaryWriter that train_dual.py:132-135,222-226,264-267 uses.
    Every scalar is also kept in `scalars.jsonl` in the same directory (readable without tensorboard)."""

    def __init__(self, log_dir, tensorboard=True):
        import json
        os.makedirs(log_dir, exist_ok=True)
        self._json, self._file, self._tb = json, open(os.path.join(log_dir, "scalars.jsonl"), "a"), None
        if tensorboard:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self._tb = SummaryWriter(log_dir)
            except Exception:                    # tensorboard is optional; the jsonl file is the record then
                self._tb = None

    def add_scalar(self, tag, value, step):
        self._file.write(self._json.dumps({"tag": tag, "value": float(value), "step": int(step)}) + "\n")
        if self._tb is not None:
            self._tb.add_scalar(tag, float(value), step)

    def add_text(self, tag, text):
        self._file.write(self._json.dumps({"tag": tag, "text": text}) + "\n")
        if self._tb is not None:
            self._tb.add_text(tag, text)

    def close(self):
        self._file.close()
        if self._tb is not None:
            self._tb.close()


def evaluate(net, eval_dataset, opt, device):
    """train_dual.py:235-267: node-count weighted means of both losses and both errors over the evaluation samples.
    The sums stay on the device; one read at the end."""
    net.eval()
    acc = torch.zeros(4, dtype=torch.float32, device=device)
    count_v = count_f = 0
    with torch.no_grad():
        for data in eval_dataset:
            data = [d.to(device) for d in data]
            y_v, y_f = data[0].y, data[1].y
            n_v, n_f = y_v.shape[0], y_f.shape[0]
            vert_p, norm_p, _ = net(data)
            acc += torch.stack((network.loss_v(vert_p, y_v, opt.loss_v) * n_v, network.loss_n(norm_p, y_f, opt.loss_n) * n_f,
                                network.error_v(vert_p, y_v) * n_v, network.error_n(norm_p, y_f) * n_f))
            count_v += n_v
            count_f += n_f
    loss_v, loss_f, error_v, error_f = (acc / torch.tensor([count_v, count_f, count_v, count_f], device=device).clamp_min(1)).tolist()
    return loss_v, loss_f, error_v, error_f


def train_epoch(net, optimizer, samples, opt, device, writer=None, first_iteration=0):
    """train_dual.py:199-230 over `samples` (this rank's share of the epoch, in order): forward, loss / batch_size, backward,
    and every `batch_size` steps (or at the end) one all-reduce + optimiser step.  With W ranks the effective batch is
    W * batch_size meshes and the loss is pre-scaled by 1 / (W * batch_size).  Returns the last step's
    (loss_v, loss_f, dual_loss, error_v, error_f) as floats."""
    _, world = parallel.rank_world()
    net.train()
    optimizer.zero_grad()
    n, last = len(samples), None
    for step, data in enumerate(samples):
        data = [d.to(device) for d in data]
        y_v, y_f = data[0].y, data[1].y
        vert_p, norm_p, _ = net(data)
        l_v = network.loss_v(vert_p, y_v, opt.loss_v)
        l_f = network.loss_n(norm_p, y_f, opt.loss_n)
        loss = network.dual_loss(l_v, l_f, v_scale=opt.loss_v_scale, n_scale=opt.loss_n_scale)
        (loss / (opt.batch_size * world)).backward()
        if (step + 1) % opt.batch_size == 0 or step + 1 == n:
            parallel.allreduce_gradients(net.parameters(), average=False)
            optimizer.step()
            optimizer.zero_grad()
            with torch.no_grad():
                last = torch.stack((l_v, l_f, loss, network.error_v(vert_p, y_v), network.error_n(norm_p, y_f))).tolist()
            if writer is not None:
                for tag, value in zip(("loss_v", "loss_f", "dual_loss", "error_v", "error_f"), last):
                    writer.add_scalar(tag, value, first_iteration + step)
    return last


def train(opt, dataset_root=None, log_root=None, tensorboard=True, net_factory=None):
    """train_dual.py:100-288.  Run directory `<log_root>/GeoBi-GNN_<data_type>_<flag>/<time>/` with training_info.txt,
    `*_params.pth`, `*_model.pth` (best validation normal error so far) and the train / test scalars.  Returns the path of the
    params file (what upstream hands to predict_dir).  Upstream's `code_bak` copy of its own sources is not made.
    `net_factory(opt) -> module` replaces DualGNN (the loop's own tests drive it with a stand-in on the CPU over a cached data
    set); DualGNN itself needs a CUDA device."""
    from . import checkpoint, dataset
    rank, world = parallel.rank_world()
    device = torch.device(f"cuda:{opt.gpu}" if (opt.gpu >= 0 and torch.cuda.is_available()) else "cpu")
    if device.type != "cuda" and net_factory is None:                     # before any run directory is made
        raise RuntimeError("training needs a CUDA device: the graph convolutions have no CPU path")
    training_name = f"GeoBi-GNN_{opt.data_type}"
    training_time = datetime.now().strftime("%Y%m%d-%H%M%S")
    if opt.seed is None:
        opt.seed = random.randint(1, 10000)
    if world > 1:                                 # one run directory and one seed (= one set of initial weights) for all ranks
        import torch.distributed as dist
        shared = [training_time, opt.seed]
        dist.broadcast_object_list(shared, src=0)
        training_time, opt.seed = shared
    flag = opt.flag
    opt.flag = f"{training_name}_{flag}_{training_time}"
    random.seed(opt.seed)
    np.random.seed(opt.seed + rank)               # the augmentation stream (RandomRotate): distinct rotations on every rank
    torch.manual_seed(opt.seed)

    log_dir = os.path.join(dataset.LOG_DIR if log_root is None else log_root, f"{training_name}_{flag}", training_time)
    os.makedirs(log_dir, exist_ok=True)
    tee = None
    if rank == 0:
        tee = sys.stdout = PrintLogger(os.path.join(log_dir, "training_info.txt"))
    try:
        print("===" * 30)
        print(f"Training flag: {opt.flag}")
        print(f"Random seed: {opt.seed} \n")
        opt.model_name = f"{training_name}_model.pth"
        opt.params_name = f"{training_name}_params.pth"
        model_name = os.path.join(log_dir, opt.model_name)
        params_name = os.path.join(log_dir, opt.params_name)
        train_writer = test_writer = None
        if rank == 0:
            checkpoint.save_params(opt, params_name)
            print(str(opt))
            train_writer = ScalarWriter(os.path.join(log_dir, "train"), tensorboard)
            test_writer = ScalarWriter(os.path.join(log_dir, "test"), tensorboard)
            test_writer.add_text("train_params", str(opt))

        if world > 1 and rank != 0:                  # rank 0 builds the cache files, the others then find them
            torch.distributed.barrier()
        list_file = lambda name: name if os.path.exists(os.path.join(dataset_root or dataset.DATASET_DIR, opt.data_type, name)) else None
        train_set = dataset.DualDataset(opt.data_type, "train", data_list_txt=list_file("train_list.txt"),
                                        filter_patch_count=opt.filter_patch_count, submesh_size=opt.sub_size,
                                        transform=dataset.RandomRotate(False), root=dataset_root, device=device)
        eval_set = dataset.DualDataset(opt.data_type, "test", data_list_txt=list_file("test_list.txt"), submesh_size=opt.sub_size,
                                       root=dataset_root, device=device)
        if world > 1 and rank == 0:
            torch.distributed.barrier()
        print(f"\nTraining set: {len(train_set):>4} samples")
        print(f"Testing set:  {len(eval_set):>4} samples")
        print("===" * 30)

        if net_factory is not None:
            net = net_factory(opt)
        else:
            net = network.DualGNN(force_depth=opt.force_depth, pool_type=opt.pool_type, wei_param=opt.wei_param)
        print(f"Total parameters: {sum(p.numel() for p in net.parameters())}")
        last_epoch = 0
        if opt.restore:
            checkpoint.load_model(net, opt.model_path)
            last_epoch = 500                                          # train_dual.py:158
        net = net.to(device)
        optimizer = make_optimizer(opt, net.parameters())
        lr_sch = make_scheduler(opt, optimizer)

        print("Start training ...")
        time_start = datetime.now()
        best_error = float("inf")
        order_rng = np.random.RandomState(opt.seed)                   # the same shuffle on every rank
        for epoch in range(last_epoch, opt.max_epoch):
            print_log = epoch % 10 == 0
            order = order_rng.permutation(len(train_set))             # DataLoader(shuffle=True)
            n_steps = len(order) // world * world if world > 1 else len(order)          # equal step counts: the all-reduces pair up
            mine = _LazySamples(train_set, order[:n_steps][rank::world])
            first_iteration = len(order) * (epoch - 1)                # as upstream counts it (train_dual.py:200)
            train_epoch(net, optimizer, mine, opt, device, train_writer, first_iteration)
            last_lr = optimizer.param_groups[0]["lr"]
            iteration = first_iteration + max(len(order) - 1, 0)
            loss_v, loss_f, error_v, error_f = evaluate(net, eval_set, opt, device)
            if test_writer is not None:
                for tag, value in (("loss_v", loss_v), ("loss_f", loss_f), ("error_v", error_v), ("error_f", error_f)):
                    test_writer.add_scalar(tag, value, iteration)
            if opt.lr_sch == "auto":
                lr_sch.step(error_f)
            else:
                lr_sch.step()
            span = datetime.now() - time_start
            str_log = (f"Epoch {epoch:>3}: {str(span).split('.')[0]:>8}  loss:{loss_v:.4f} {loss_f:.4f} | "
                       f"error:{error_v:.4f} {error_f:.4f}  lr:{last_lr:.4e}")
            if error_f < best_error:
                best_error = error_f
                if rank == 0:
                    checkpoint.save_model(net, model_name)
                str_log += " - save model"
                print_log = True
            if print_log and rank == 0:
                print(str_log)
        if rank == 0:
            train_writer.close()
            test_writer.close()
        print(f"\n{opt.flag}\nbest error: {best_error}")
        print("===" * 30)
        return params_name
    finally:
        if tee is not None:
            sys.stdout = tee.close()


class _LazySamples:
    """This rank's samples of one epoch, loaded (and augmented) one at a time as the loop reaches them."""

    def __init__(self, data_set, indices):
        self.data_set, self.indices = data_set, list(indices)

    def __len__(self):
        return len(self.indices)

    def __iter__(self):
        return (self.data_set[int(i)] for i in self.indices)


def main(argv=None):
    """`python -m geobi_gnn_b200.train --data_type=Synthetic --flag=... --gpu=0` = train_dual.py's __main__: train, then
    predict_dir over the test split with the new run files."""
    from . import inference
    opt = parse_arguments(argv)
    params_file = train(opt)
    print("\n--- Training end ---")
    inference.predict_dir(params_file, data_dir=None, sub_size=opt.sub_size, gpu=opt.gpu)


if __name__ == "__main__":
    main()
