"""The callback *protocol* of the reference (qucumber/callbacks/callback.py:16) — hooks only.
QuCumber's own callback classes stay in QuCumber; ``fit`` accepts any object with these methods."""
import time


class CallbackBase:
    def on_train_start(self, nn_state):
        pass

    def on_train_end(self, nn_state):
        pass

    def on_epoch_start(self, nn_state, epoch):
        pass

    def on_epoch_end(self, nn_state, epoch):
        pass

    def on_batch_start(self, nn_state, epoch, batch):
        pass

    def on_batch_end(self, nn_state, epoch, batch):
        pass


class Timer(CallbackBase):
    """Wall time between on_train_start and on_train_end (reference: qucumber/callbacks/timer.py:21-62)."""

    def __init__(self, verbose=True):
        self.verbose = verbose
        self.start_time = None
        self.total_time = None

    def on_train_start(self, nn_state):
        self.start_time = time.time()

    def on_train_end(self, nn_state):
        self.total_time = time.time() - self.start_time
        if self.verbose:
            print(f"Total time elapsed during training: {self.total_time:.3f} s")
