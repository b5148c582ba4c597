"""Filter-kernel geometry sweep: the same 1 GiB config-2 scan with libraries built with other CTA sizes / loads in flight
(make OUT=build/libacb200_f<threads>_u<unroll>.so EXTRA="-DACB_FILTER_THREADS=.. -DACB_FILTER_UNROLL=..")."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from php_aho_corasick_b200 import native
if len(sys.argv) > 1:
    native.LIB_PATH = os.path.join(ROOT, sys.argv[1])
import torch
from php_aho_corasick_b200 import workloads as W
from php_aho_corasick_b200.native import Automaton
from r02_probe import split

needles, _ = W.cfg2_needles()
a = Automaton(0); a.add_php_order(needles); a.finalize()

d = torch.from_numpy(W.cfg2_stream(0, 0, 512)).cuda()
split(f"{sys.argv[1] if len(sys.argv) > 1 else 'default':40s}", a, lambda: a.search_device_uniform(d.data_ptr(), 512 * 256, 8192)[1], reps=8)
