import cProfile, pstats, sys, os, io
sys.argv = [sys.argv[0]]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "scripts"))
import importlib.util
spec = importlib.util.spec_from_file_location("det", os.path.join(ROOT, "scripts", "dropin_episode_time.py"))
src = open(os.path.join(ROOT, "scripts", "dropin_episode_time.py")).read().split("for name, fn in")[0]
g = {"__file__": os.path.join(ROOT, "scripts", "dropin_episode_time.py"), "__name__": "det"}
exec(compile(src, "det", "exec"), g)
g["macro_or_hybrid"]("macro", 1)
pr = cProfile.Profile(); pr.enable(); g["macro_or_hybrid"]("macro", 2); pr.disable()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45); print(s.getvalue()[:9000])
