class Checkpointer:
    def __init__(self, model, save_dir="", *, save_to_disk=True, **checkpointables):
        self.model = model
        self.checkpointables = dict(checkpointables)
        self.save_dir = save_dir
        self.save_to_disk = save_to_disk


class PeriodicCheckpointer:
    def __init__(self, checkpointer, period, max_iter=None, max_to_keep=None, file_prefix="model"):
        self.checkpointer, self.period, self.max_iter = checkpointer, int(period), max_iter


def get_missing_parameters_message(keys): return str(keys)
def get_unexpected_parameters_message(keys): return str(keys)
def _strip_prefix_if_present(state_dict, prefix): pass
class _IncompatibleKeys: pass
