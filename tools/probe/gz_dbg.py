import gzip, os, sys
sys.path.insert(0, "/root/repo")
os.environ["FRB_GZ_VERBOSE"] = "1"
from frender_b200.engine import Context
ctx = Context(0, table_log2=12)
data = (b"F" * 5000 + b"\n") * 300
open("/tmp/runs.gz", "wb").write(gzip.compress(data, 6, mtime=0))
got = ctx.gz_inflate("/tmp/runs.gz", len(data) + 1024)
print("runs:", None if got is None else got == data)
