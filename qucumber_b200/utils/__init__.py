"""Small host-side helpers mirroring qucumber.utils (reference: qucumber/utils/__init__.py:20-45)."""
import functools


class auto_unsqueeze_args:
    """Decorator: positional tensor arguments listed in ``positions`` (default: the first after
    ``self``) may be 1-D; they are promoted to a batch of one and the result loses that batch
    axis again — the reference's calling convention for single states."""

    def __init__(self, *positions):
        self.positions = tuple(positions) or (1,)

    def __call__(self, fn):
        @functools.wraps(fn)
        def wrapper(*args, **kwargs):
            args = list(args)
            promoted = False
            for pos in self.positions:
                if args[pos].dim() < 2:
                    args[pos] = args[pos].unsqueeze(0)
                    promoted = True
            result = fn(*args, **kwargs)
            return result.squeeze_(0) if promoted else result

        return wrapper
