import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import tune_deform as t
for f in (1, 8, 64, 400, 1700):
    t.run("train", f, [None])
