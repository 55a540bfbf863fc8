"""`gaussian_splatting.utils.system_utils.mkdir_p` (imported by the reference's evaluation / saving code) [UPSTREAM-RECALL]."""
from errno import EEXIST
from os import makedirs, path


def mkdir_p(folder_path):
    try:
        makedirs(folder_path)
    except OSError as exc:
        if exc.errno == EEXIST and path.isdir(folder_path):
            pass
        else:
            raise
