"""`ModelCheckpoint` and `EarlyStopping` as data_utils.py:16-35 constructs them (monitor / mode / patience)."""
import os

import torch


class _Monitor:
    def __init__(self, monitor="val_loss", mode="min"):
        self.monitor, self.mode = monitor, mode
        self.best = None

    def _improved(self, value):
        return self.best is None or (value < self.best if self.mode == "min" else value > self.best)


class ModelCheckpoint(_Monitor):
    def __init__(self, monitor="val_loss", dirpath="checkpoints", filename="model", mode="min", **kwargs):
        super().__init__(monitor, mode)
        self.dirpath, self.filename = dirpath, filename
        self.best_model_path = ""
        self.best_model_score = None

    def on_validation_end(self, trainer, model):
        value = trainer.callback_metrics.get(self.monitor)
        if value is None or not self._improved(value):
            return
        self.best = self.best_model_score = value
        os.makedirs(self.dirpath, exist_ok=True)
        self.best_model_path = os.path.join(self.dirpath, self.filename + ".ckpt")
        torch.save({"state_dict": {k: v.detach().cpu() for k, v in model.state_dict().items()}, "epoch": trainer.current_epoch,
                    "best_model_score": value}, self.best_model_path)


class EarlyStopping(_Monitor):
    def __init__(self, monitor="val_loss", patience=3, verbose=False, mode="min", **kwargs):
        super().__init__(monitor, mode)
        self.patience, self.verbose, self.wait = int(patience), verbose, 0

    def on_validation_end(self, trainer, model):
        value = trainer.callback_metrics.get(self.monitor)
        if value is None:
            return
        if self._improved(value):
            self.best, self.wait = value, 0
        else:
            self.wait += 1
            if self.wait >= self.patience:
                if self.verbose:
                    print(f"[shim trainer] early stop: {self.monitor} did not improve for {self.patience} epochs")
                trainer.should_stop = True
