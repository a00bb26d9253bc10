"""Helper (not a test): train one of the reference's UNMODIFIED task models (oracle/_ref/models/{planetoid,ppi,pattern}_gat.py
on top of GATModel.py) for a few epochs and print the loss curve.

    python tests/_train_curve.py {b200|reference} <Cora|Pubmed|PPI|PATTERN> <epochs> [key=value ...]

Same construction as the reference's train.py:run (config from run_config.data_config, `Model(**config)`, `Trainer.fit`),
with `pl.seed_everything(42)` as the reference notebook does, dropout forced to 0 (Philox cannot bit-match nn.Dropout, so a
curve comparison needs it off) and planted labels (the stand-in datasets).  `b200` = overlay in front (B200 layer),
`reference` = the reference layer, both on the CUDA device the stand-in Trainer picks.  Prints `CURVE {...}`."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import _run_ref_script as R


def main():
    layer, dataset, epochs, extra = sys.argv[1], sys.argv[2], int(sys.argv[3]), sys.argv[4:]
    R.setup_path(layer)
    import pytorch_lightning as pl
    import torch
    from run_config import data_config
    from models.pattern_gat import PatternGAT
    from models.planetoid_gat import PlanetoidGAT
    from models.ppi_gat import PPI_GAT

    config = dict(data_config[dataset], dataset=dataset, num_epochs=epochs, exec_type="train", dropout=0.0)
    for kv in extra:
        k, v = kv.split("=")
        config[k] = float(v)
    pl.seed_everything(42)
    cls = {"PPI": PPI_GAT, "PATTERN": PatternGAT}.get(dataset, PlanetoidGAT)
    model = cls(**config)
    trainer = pl.Trainer(max_epochs=epochs, callbacks=[])
    trainer.fit(model)
    import models.gat_layer as layer_mod
    launches = 0
    if layer == "b200":
        from gat_pytorch_b200 import _lib
        launches = int(_lib.load().gat_launch_count())
    print("CURVE " + json.dumps({"layer_module": layer_mod.GATLayer.__module__, "device": str(trainer.device), "launches": launches,
                                 "step_losses": trainer.step_losses, "history": trainer.history,
                                 "cuda": torch.cuda.is_available()}), flush=True)


if __name__ == "__main__":
    main()
