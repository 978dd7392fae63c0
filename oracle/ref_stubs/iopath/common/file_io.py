import os
from contextlib import contextmanager


class PathHandler:
    pass


class HTTPURLHandler(PathHandler):
    pass


class OneDrivePathHandler(PathHandler):
    pass


class PathManager:
    def register_handler(self, handler, allow_override=False): pass
    def open(self, path, mode="r", **k): return open(path, mode)
    def isfile(self, p): return os.path.isfile(p)
    def isdir(self, p): return os.path.isdir(p)
    def exists(self, p): return os.path.exists(p)
    def ls(self, p): return os.listdir(p)
    def mkdirs(self, p): os.makedirs(p, exist_ok=True)
    def get_local_path(self, p, **k): return p
    def rm(self, p): os.remove(p)
    def copy(self, a, b, overwrite=False):
        import shutil
        shutil.copy(a, b)
        return True


@contextmanager
def file_lock(path):
    yield
