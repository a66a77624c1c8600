import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cProfile, pstats, io, time, os, tempfile
import workloads
from squarna_b200 import SQUARNA as CLI
rows, ref = workloads.config4(2000, 300, 400)
path = "/tmp/c4.afa"
open(path, "w").write(workloads.config4_text(rows, ref))
buf = io.StringIO(); CLI.Predict(inputfile=path, alignment=True, write_to=buf)
pr = cProfile.Profile(); pr.enable()
t = time.perf_counter(); buf = io.StringIO(); CLI.Predict(inputfile=path, alignment=True, write_to=buf); print("total", time.perf_counter() - t)
pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
