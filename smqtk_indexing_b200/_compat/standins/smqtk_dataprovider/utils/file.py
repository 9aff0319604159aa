"""Stand-in for ``smqtk_dataprovider.utils.file``."""
import errno
import os


def safe_create_dir(d: str) -> str:
    d = os.path.abspath(os.path.expanduser(d))
    try:
        os.makedirs(d)
    except OSError as ex:
        if ex.errno == errno.EEXIST and os.path.exists(d):
            pass
        else:
            raise
    return d
