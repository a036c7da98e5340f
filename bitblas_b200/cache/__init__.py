"""``bitblas.cache`` names kept for callers (bitblas/cache/operator.py:24-135; used by module/__init__.py:16,245-256 and
integration/BitNet/utils_quant.py:9).  There is nothing to persist: operators are descriptors over a prebuilt
library, so the "database" is an in-memory dict and save/load are no-ops."""
import os
import threading

BITBLAS_DATABASE_PATH = os.path.expanduser(os.environ.get("BITBLAS_DEFAULT_CACHE_PATH", "~/.cache/bitblas"))


class OperatorCache:
    cache_locker = threading.RLock()

    def __init__(self):
        self.cache = {}

    def add(self, config, op_inst):
        with self.cache_locker:
            self.cache[config] = op_inst

    def get(self, config):
        with self.cache_locker:
            return self.cache.get(config)

    def exists(self, config):
        return config in self.cache

    def clear(self):
        with self.cache_locker:
            self.cache.clear()

    def size(self):
        return len(self.cache)

    def save_into_database(self, database_path=None, target=None):
        return None

    def load_from_database(self, database_path, target=None):
        return None


global_operator_cache = OperatorCache()


def load_global_ops_cache(database_path=None, target=None):
    return global_operator_cache


def get_database_path():
    return BITBLAS_DATABASE_PATH


def set_database_path(path):
    global BITBLAS_DATABASE_PATH
    BITBLAS_DATABASE_PATH = path
    return BITBLAS_DATABASE_PATH
