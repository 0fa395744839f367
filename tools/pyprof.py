"""Where does the host time of a small QAT step go?  cProfile over YOLOv8n b2@320 steps."""
import cProfile, pstats, sys, os, io
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from benchmarks import yolo_qat
args = yolo_qat.parse(["--model", "n", "--batch", "2", "--imgsz", "320", "--steps", "30"])
pr = cProfile.Profile()
orig = yolo_qat.run
res = None
def wrapped(a):
    return orig(a)
pr.enable(); res = wrapped(args); pr.disable()
print(res)
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
