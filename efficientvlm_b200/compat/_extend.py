"""Shared helper for the drop-in packages: append the reference's same-named package directory to `__path__` so that
submodules we do not replace (task models such as models/model_pretrain.py) keep resolving to the reference's files."""
import os
import sys


def extend_path(pkg_name, pkg_path):
    here = os.path.dirname(os.path.abspath(pkg_path[0]))
    roots = [os.environ.get("EVLM_REFERENCE_ROOT", "")] + list(sys.path)
    for root in roots:
        if not root:
            continue
        cand = os.path.join(os.path.abspath(root), pkg_name)
        if os.path.isdir(cand) and os.path.abspath(cand) != os.path.abspath(pkg_path[0]) and cand not in pkg_path:
            if os.path.abspath(root) == here:
                continue
            pkg_path.append(cand)
            return cand
    return None
