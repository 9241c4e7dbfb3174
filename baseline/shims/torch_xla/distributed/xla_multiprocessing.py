"""xmp.spawn: processes are already created by torchrun, so run the function in-process."""
import os


def spawn(fn, args=(), nprocs=None, **kwargs):
    fn(int(os.environ.get("LOCAL_RANK", 0)), *args)
