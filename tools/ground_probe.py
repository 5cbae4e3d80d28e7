import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')); sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import redmax_b200 as rb
from config_bench import fwd
for gz, h, label in ((-1e6, 1e-3, 'ground far away (no contact) h=1e-3 BDF1'), (-40.0, 2e-4, 'ground z=-40 h=2e-4 BDF2'), (-40.0, 1e-4, 'ground z=-40 h=1e-4 BDF2')):
    sc = rb.chain_scene(32, ground=True, h=h, nsteps=100, ground_z=gz); sc.init()
    fwd(label, sc, 4096, 1 if gz < -1e5 else 2, 100, reps=2)
