"""Classification metrics with the reference's interface (utils/metrics.py: Metric / Accuracy / Precision / Recall /
F1: ``update((y_pred, batch))``, ``compute()``, ``get_output()``, ``set_device``, ``sync_across_processes``).

What changes: the counters live on the scoring device and ``update`` never calls ``.item()`` (utils/metrics.py:53 syncs
the host once per batch, which serialises launch sequences once ``classify`` is fast); ``sync_across_processes`` sums
the counters with ONE ``torch.distributed.all_reduce`` per metric (an ``accelerate.Accelerator`` is accepted and used
when given, as in the reference)."""
import torch


class Metric(torch.nn.Module):
    _fields = ()

    def __init__(self, name, device=torch.device("cpu")):
        super().__init__()
        self.name = name
        self.device = device
        self.required_output_keys = ()
        self.reset()

    def reset(self):
        self._c = torch.zeros(max(len(self._fields), 1), dtype=torch.int64, device=self.device)

    def set_device(self, device):
        self.device = device
        self._c = self._c.to(device)

    def _add(self, *vals):
        self._c += torch.stack([v.sum() for v in vals]).to(self._c.device)

    def update(self, output):
        pass

    def compute(self):
        pass

    def get_output(self, reduce=True):
        return self.compute()

    def sync_across_processes(self, accelerator=None):
        if accelerator is not None:
            self._c = accelerator.reduce(self._c)
            return
        import torch.distributed as dist
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            dist.all_reduce(self._c)

    def __call__(self, output):
        self.update(output)
        return self.compute()

    @staticmethod
    def _ratio(num, den):
        return 0.0 if float(den) == 0 else num.float() / den.float()


class Accuracy(Metric):
    _fields = ("correct", "total")

    def update(self, output):
        y_pred, batch = output
        y_true = batch['prompt'].to(y_pred.device)
        self._add((y_pred == y_true), torch.ones_like(y_true, dtype=torch.bool))

    correct = property(lambda self: self._c[0])
    total = property(lambda self: self._c[1])

    def compute(self):
        return {self.name: self.correct / self.total}


class _Binary(Metric):
    _fields = ("tp", "fp", "fn")

    def __init__(self, name=None, device=torch.device("cpu")):
        super().__init__(name or type(self).__name__.lower(), device)

    def update(self, output):
        y_pred, batch = output
        y_true = batch["prompt"].to(y_pred.device)
        self._add((y_pred == 1) & (y_true == 1), (y_pred == 1) & (y_true == 0), (y_pred == 0) & (y_true == 1))

    tp = property(lambda self: self._c[0])
    fp = property(lambda self: self._c[1])
    fn = property(lambda self: self._c[2])


class Precision(_Binary):
    def compute(self):
        return {self.name: self._ratio(self.tp, self.tp + self.fp)}


class Recall(_Binary):
    def compute(self):
        return {self.name: self._ratio(self.tp, self.tp + self.fn)}


class F1(_Binary):
    def compute(self):
        return {self.name: self._ratio(2 * self.tp, 2 * self.tp + self.fp + self.fn)}
