"""`TensorBoardLogger` is only a type annotation and an optional histogram sink in the reference (GATModel.py:214-220, :236-253);
this stand-in swallows the calls."""


class _Experiment:
    def add_histogram(self, *args, **kwargs):
        pass

    def add_scalar(self, *args, **kwargs):
        pass


class TensorBoardLogger:
    def __init__(self, save_dir="lightning_logs", name="default", **kwargs):
        self.save_dir, self.name = save_dir, name
        self.experiment = _Experiment()
