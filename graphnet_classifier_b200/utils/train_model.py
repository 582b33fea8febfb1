"""``train`` with the reference's signature, loop order, checkpoint names and log
format (reference utils/train_model.py:8-81): Adam(1e-3) + cross entropy, one
optimizer step per item of ``dataset``, early stopping on the *training* loss,
``best_model_epoch{n}.pth`` / ``final_model.pth`` and a timestamped text log.

Extensions that leave the single-process behaviour unchanged:
  * labels - and plain tensor samples from a host loader - are moved to the model's device (the reference is
    CPU-only);
  * ``grad_sync`` - a callable run between ``backward()`` and ``step()``; the
    data-parallel launcher passes ``GradBucket.all_reduce`` (utils/distributed.py);
  * only rank 0 writes files when ``torch.distributed`` is initialised; the epoch loss that drives the best-model /
    early-stopping decisions is averaged over all ranks, so every rank leaves the loop in the same epoch;
  * ``cuda_graph`` (default: on for CUDA models without ``grad_sync``) - the reference makes one optimizer step per
    graph, so a step is ~350 small launches; the whole step (forward, loss, backward, Adam) is captured once per
    sample shape as a CUDA graph and replayed with the item's tensors copied into its static inputs.  Same kernels
    and same update rule; items whose shapes or topology differ from the captured ones run eagerly.
"""
from __future__ import annotations

import os
import time
from datetime import datetime

import torch
import torch.nn as nn
import torch.optim as optim


def _is_rank0() -> bool:
    d = torch.distributed
    return not (d.is_available() and d.is_initialized()) or d.get_rank() == 0


def _global_mean_loss(running: float, steps: int, device) -> float:
    """Epoch loss over ALL ranks' shards.  Every rank must take the same best-model / early-stopping decision
    (reference utils/train_model.py:57-69): a rank that left the loop on its own shard's loss would leave the others
    blocked in the next gradient all-reduce."""
    d = torch.distributed
    if not (d.is_available() and d.is_initialized()) or d.get_world_size() == 1:
        return running / max(1, steps)
    on_gpu = d.get_backend() == "nccl"
    t = torch.tensor([running, float(steps)], dtype=torch.float64, device=device if on_gpu else "cpu")
    d.all_reduce(t)
    return float(t[0]) / max(1.0, float(t[1]))


class _GraphedStep:
    """One training step (forward, cross entropy, backward, Adam) captured as a CUDA graph for one sample layout:
    a tuple of CUDA tensors whose ``edge_index`` carries a prebuilt topology (``ops.attach_graph``; the cached grid
    of the pixel / patch builders), or a single CUDA tensor."""

    def __init__(self, model, optimizer, criterion, sample, label):
        self.model, self.optimizer, self.criterion = model, optimizer, criterion
        self.tuple_in = isinstance(sample, (tuple, list))
        parts = tuple(sample) if self.tuple_in else (sample,)
        self.key = self.layout(sample, label)
        # floating-point inputs change per item and are copied in; index tensors (edge_index) are topology:
        # the captured kernels read the attached CSR, so an item must carry the very same topology object.  The
        # positions the builder emitted with that topology are part of it (same tensor for every item; the model's
        # edge-class shortcut recognises it by identity), so they are kept, not cloned
        self.fixed = self._topology_bound(parts)
        self.static = tuple(p if fx else p.clone() for p, fx in zip(parts, self.fixed))
        self.label = label.clone()
        self.loss = None
        self.graph = None

    @staticmethod
    def _topology_bound(parts):
        topos = [getattr(p, "_gnc_graph", None) for p in parts if not p.is_floating_point()]
        return tuple((not p.is_floating_point()) or any(t is not None and t.classes_valid_for(p) for t in topos) for p in parts)

    @staticmethod
    def layout(sample, label):
        parts = tuple(sample) if isinstance(sample, (tuple, list)) else (sample,)
        if not all(isinstance(p, torch.Tensor) and p.is_cuda for p in parts) or not isinstance(label, torch.Tensor):
            return None
        key = []
        fixed = _GraphedStep._topology_bound(parts)
        for p, fx in zip(parts, fixed):
            if p.is_floating_point():
                key.append((tuple(p.shape), p.dtype, id(p) if fx else None))
            else:
                topo = getattr(p, "_gnc_graph", None)
                if topo is None:
                    return None
                key.append((tuple(p.shape), id(topo)))
        return tuple(key) + (tuple(label.shape),)

    def _step(self):
        sample = self.static if self.tuple_in else self.static[0]
        loss = self.criterion(self.model(sample), self.label)
        self.optimizer.zero_grad(set_to_none=True)
        loss.backward()
        self.optimizer.step()
        return loss.detach()

    def capture(self):
        dev = self.label.device
        side = torch.cuda.Stream(device=dev)
        # the warm-up steps (kernel attributes, workspaces, allocator pools) must not train: parameters and the
        # optimizer state are snapshotted and put back
        snap = [p.detach().clone() for p in self.model.parameters()]
        opt_snap = {p: {k: v.clone() for k, v in st.items() if isinstance(v, torch.Tensor)}
                    for p, st in self.optimizer.state.items()}
        grads = [p.grad for p in self.model.parameters()]
        side.wait_stream(torch.cuda.current_stream(dev))     # after the snapshots have been enqueued
        with torch.cuda.stream(side):
            for _ in range(2):
                self._step()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        with torch.no_grad():
            for p, s0 in zip(self.model.parameters(), snap):
                p.copy_(s0)
            for p, st in self.optimizer.state.items():
                before = opt_snap.get(p, {})
                for k, v in st.items():
                    if isinstance(v, torch.Tensor):
                        v.copy_(before[k]) if k in before else v.zero_()     # no state before = a fresh optimizer
        for p, g in zip(self.model.parameters(), grads):
            p.grad = g
        del snap, opt_snap, grads
        self.graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.graph):
            self.loss = self._step()
        # the capture itself does not execute: parameters and optimizer state are still at their start values

    def run(self, sample, label):
        parts = tuple(sample) if self.tuple_in else (sample,)
        for dst, src, fx in zip(self.static, parts, self.fixed):
            if not fx:
                dst.copy_(src, non_blocking=True)
        self.label.copy_(label, non_blocking=True)
        self.graph.replay()
        return self.loss


def train(model, dataset, epochs, patience=5, output_path="weights", start_weights=None, grad_sync=None,
          cuda_graph=None):
    params = list(model.parameters())
    on_cuda = bool(params) and all(p.is_cuda for p in params)
    if cuda_graph is None:
        cuda_graph = on_cuda and grad_sync is None
    # capturable: Adam keeps its step count on the device, which a replayed graph needs; the update rule is the same
    optimizer = optim.Adam(model.parameters(), lr=1e-3, capturable=True) if (cuda_graph and on_cuda) else \
        optim.Adam(model.parameters(), lr=1e-3)
    criterion = nn.CrossEntropyLoss()
    graphed = {}                                         # sample layout -> _GraphedStep
    best_loss = float("inf")
    stale_epochs = 0
    rank0 = _is_rank0()

    if start_weights:
        model.load_state_dict(torch.load(start_weights))

    log_path = None
    if rank0:
        os.makedirs(output_path, exist_ok=True)
        print(f"Training model in {output_path}")
        stamp = datetime.now().strftime("%Y%m%d_%H%M%S")
        log_path = os.path.join(output_path, f"training_logs_{stamp}.txt")
        with open(log_path, "w") as log:
            log.write(f"Training started at: {datetime.now().strftime('%Y-%m-%d %H:%M:%S')}\n")
            log.write(f"Epochs: {epochs}, Patience: {patience}\n")
            log.write(f"Output path: {output_path}\n")
            log.write("-" * 50 + "\n")

    for epoch in range(epochs):
        t0 = time.time()
        running, steps = 0.0, 0
        for sample, label in dataset:
            if on_cuda and isinstance(sample, torch.Tensor) and not sample.is_cuda:
                sample = sample.to(params[0].device, non_blocking=True)     # host loader batches (MLP baseline)
            key = None
            if cuda_graph and on_cuda:
                if isinstance(label, torch.Tensor) and not label.is_cuda:
                    label = label.to(params[0].device)
                key = _GraphedStep.layout(sample, label)
            if key is not None:
                step = graphed.get(key)
                if step is None:
                    if len(graphed) >= 4:                # a handful of layouts at most (pixel graphs have one)
                        graphed.clear()
                    step = graphed[key] = _GraphedStep(model, optimizer, criterion, sample, label)
                    step.capture()
                loss = step.run(sample, label)
            else:
                logits = model(sample)                  # tensor for MLP, (x, pos, edge_index) for GNN
                if isinstance(label, torch.Tensor) and label.device != logits.device:
                    label = label.to(logits.device)
                loss = criterion(logits, label)
                optimizer.zero_grad()
                loss.backward()
                if grad_sync is not None:
                    grad_sync()
                optimizer.step()
            running += loss.item()
            steps += 1
        avg_loss = _global_mean_loss(running, steps, params[0].device if params else "cpu")
        t1 = time.time()
        if rank0:
            print(f"Epoch {epoch+1}/{epochs}, avg_loss={avg_loss:.4f}")
            print(f"epoch: {epoch + 1} needed {t1 - t0} time")
            with open(log_path, "a") as log:
                log.write(f"Epoch {epoch+1}/{epochs}, avg_loss={avg_loss:.4f}\n")
                log.write(f"Epoch {epoch+1}/{epochs}, needed {(t1 - t0) / 60:.2f} minutes\n")
        if avg_loss < best_loss:
            best_loss, stale_epochs = avg_loss, 0
            if rank0:
                best_path = os.path.join(output_path, f"best_model_epoch{epoch+1}.pth")
                torch.save(model.state_dict(), best_path)
                print(f"Saved best model: {best_path}")
        else:
            stale_epochs += 1
        if stale_epochs >= patience:
            if rank0:
                print(f"Early stopping at epoch {epoch+1}")
            break

    final_path = os.path.join(output_path, "final_model.pth")
    if rank0:
        torch.save(model.state_dict(), final_path)
        print(f"Saved final model: {final_path}")
        with open(log_path, "a") as log:
            log.write("-" * 50 + "\n")
            log.write(f"Training completed at: {datetime.now().strftime('%Y-%m-%d %H:%M:%S')}\n")
            log.write(f"Best loss achieved: {best_loss:.4f}\n")
            log.write(f"Final model saved: {final_path}\n")
    return best_loss
