import sys

from . import convert

if len(sys.argv) != 3:
    raise SystemExit('usage: python -m vfs_b200.convert SRC DST')
convert(sys.argv[1], sys.argv[2])
