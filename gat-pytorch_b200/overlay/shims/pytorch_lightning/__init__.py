"""Offline stand-in for the slice of `pytorch_lightning` (pinned 1.2.3 by the reference, env/gat_req_mac_version.yml:169)
that loodvn/gat-pytorch uses -- SURVEY.md section 8-f2.  NOT the product path: it exists so that the reference's own
`train.py` / `planetoid_gat.py` / `ppi_gat.py` / `pattern_gat.py` run UNCHANGED in an image that has neither Lightning nor a
network, on top of either the reference layer (CPU) or the B200 layer (overlay).  Put this directory on PYTHONPATH only
when the real package is absent.

Surface (what the reference touches): `LightningModule` (`log`, `logger`, `device`, `load_from_checkpoint`, the
`prepare_data` / `*_dataloader` / `configure_optimizers` / `*_step` / `on_after_backward` hooks), `Trainer(max_epochs,
callbacks, ...)` with `.fit` / `.test`, `seed_everything`, `callbacks.ModelCheckpoint` / `EarlyStopping`,
`loggers.TensorBoardLogger`.  The loops are the plain ones: one optimiser, no accumulation, validation after every epoch.
"""
from __future__ import annotations

import os
import random

import numpy as np
import torch
from torch import nn

from . import callbacks, loggers  # noqa: F401

__version__ = "0.0-gat-b200-shim"


def seed_everything(seed: int = 42) -> int:
    random.seed(seed)
    np.random.seed(seed)
    torch.manual_seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)
    return seed


class LightningModule(nn.Module):
    def __init__(self, *args, **kwargs):
        super().__init__()
        self.logger = None
        self.trainer = None
        self._logged = {}

    # -- what the reference's models call -------------------------------------------------------------------------
    @property
    def device(self):
        for p in self.parameters():
            return p.device
        return torch.device("cpu")

    def log(self, name, value, *args, **kwargs):
        if torch.is_tensor(value):
            value = value.detach().float().mean().item()
        self._logged.setdefault(name, []).append(float(value))

    @classmethod
    def load_from_checkpoint(cls, checkpoint_path, map_location=None, strict=True, **kwargs):
        ckpt = torch.load(checkpoint_path, map_location=map_location or "cpu", weights_only=False)
        model = cls(**kwargs)
        model.load_state_dict(ckpt["state_dict"] if "state_dict" in ckpt else ckpt, strict=strict)
        return model

    # -- hooks (defaults) -----------------------------------------------------------------------------------------
    def prepare_data(self):
        pass

    def on_after_backward(self):
        pass

    def configure_optimizers(self):
        raise NotImplementedError

    def val_dataloader(self):
        return None

    def test_dataloader(self):
        return None


def _to_device(batch, device):
    return batch.to(device) if hasattr(batch, "to") else batch


class Trainer:
    def __init__(self, max_epochs=1000, callbacks=None, fast_dev_run=False, gpus=None, logger=None, **kwargs):
        self.max_epochs = 1 if fast_dev_run else int(max_epochs)
        self.callbacks = list(callbacks or [])
        self.logger = logger if logger not in (True, False) else None
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.model = None
        self.current_epoch = 0
        self.should_stop = False
        self.callback_metrics = {}
        self.history = []           # per-epoch metric dicts (what ModelCheckpoint / EarlyStopping saw)
        self.step_losses = []       # training loss of every optimiser step (training-parity tests compare two runs)

    def _epoch_metrics(self, model):
        out = {k: float(np.mean(v)) for k, v in model._logged.items() if v}
        model._logged = {}
        return out

    def fit(self, model):
        self.model = model
        model.trainer, model.logger = self, self.logger
        model.prepare_data()
        model.to(self.device)
        conf = model.configure_optimizers()
        scheduler, monitor = None, "val_loss"
        if isinstance(conf, dict):
            optimizer, scheduler, monitor = conf["optimizer"], conf.get("lr_scheduler"), conf.get("monitor", "val_loss")
        else:
            optimizer = conf[0] if isinstance(conf, (list, tuple)) else conf
        for epoch in range(self.max_epochs):
            self.current_epoch = epoch
            model.train()
            for i, batch in enumerate(model.train_dataloader()):
                loss = model.training_step(_to_device(batch, self.device), i)
                self.step_losses.append(float(loss.detach()))
                optimizer.zero_grad()
                loss.backward()
                model.on_after_backward()
                optimizer.step()
            loader = model.val_dataloader()
            if loader is not None:
                model.eval()
                with torch.no_grad():
                    for i, batch in enumerate(loader):
                        model.validation_step(_to_device(batch, self.device), i)
            self.callback_metrics = self._epoch_metrics(model)
            self.history.append(dict(self.callback_metrics))
            if scheduler is not None and monitor in self.callback_metrics:
                scheduler.step(self.callback_metrics[monitor])
            print(f"[shim trainer] epoch {epoch}: " + ", ".join(f"{k}={v:.4f}" for k, v in sorted(self.callback_metrics.items())), flush=True)
            for cb in self.callbacks:
                cb.on_validation_end(self, model)
            if self.should_stop:
                break
        return None

    def test(self, model=None, **kwargs):
        model = model if model is not None else self.model
        if model is None:
            raise RuntimeError("Trainer.test() needs a model (none was fitted)")
        model.trainer, model.logger = self, self.logger
        if getattr(model, "test_ds", None) is None:
            model.prepare_data()
        model.to(self.device)
        model.eval()
        with torch.no_grad():
            for i, batch in enumerate(model.test_dataloader()):
                model.test_step(_to_device(batch, self.device), i)
        metrics = self._epoch_metrics(model)
        print("[shim trainer] test: " + ", ".join(f"{k}={v:.4f}" for k, v in sorted(metrics.items())), flush=True)
        return [metrics]
