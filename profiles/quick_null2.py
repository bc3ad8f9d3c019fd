"""quick_null.py with null_impl 1 vs 2 (blocks automatic) for one shape: N, L, K from the environment"""
import os
os.environ["KEY"] = "null_impl"; os.environ.setdefault("SB", "1,2")
exec(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "quick_null.py")).read())
