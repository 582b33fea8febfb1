"""``train`` with the reference's signature, loop order, checkpoint names and log
format (reference utils/train_model.py:8-81): Adam(1e-3) + cross entropy, one
optimizer step per item of ``dataset``, early stopping on the *training* loss,
``best_model_epoch{n}.pth`` / ``final_model.pth`` and a timestamped text log.

Extensions that leave the single-process behaviour unchanged:
  * labels are moved to the logits' device (the reference is CPU-only);
  * ``grad_sync`` - a callable run between ``backward()`` and ``step()``; the
    data-parallel launcher passes ``GradBucket.all_reduce`` (utils/distributed.py);
  * only rank 0 writes files when ``torch.distributed`` is initialised.
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


def train(model, dataset, epochs, patience=5, output_path="weights", start_weights=None, grad_sync=None):
    optimizer = optim.Adam(model.parameters(), lr=1e-3)
    criterion = nn.CrossEntropyLoss()
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
            logits = model(sample)                      # tensor for MLP, (x, pos, edge_index) for GNN
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
        avg_loss = running / max(1, steps)
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
