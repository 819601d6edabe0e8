"""Locate the compiled extension next to the package.

Mirrors util/icqSharedLibraryUtils.py:13-24 of the reference (resolve the
package-relative library name, glob for a suffixed file name, return the first
match) without pkg_resources.  The library name for this package is
``libicqLaplaceMatricesB200`` (the reference's was ``icqLaplaceMatricesCpp``).
"""
import glob
import os

PACKAGE_DIR = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def getFullyQualifiedSharedLibraryName(libName):
    paths = sorted(glob.glob(libName + '*.so'))
    if not paths:
        raise RuntimeError(
            'ERROR: shared library {0}*.so not found; build it with '
            '`python -c "import __graft_entry__ as g; g.build()"` or '
            '`make -C icqsol_b200/csrc` (there is no CPU fallback)'.format(libName))
    # return the first path
    return paths[0]


def getSharedLibraryName(libName):
    libName = os.path.join(PACKAGE_DIR, libName)
    return getFullyQualifiedSharedLibraryName(libName)
