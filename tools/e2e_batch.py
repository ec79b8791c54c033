"""End-to-end time of agatha_align_job on one GPU as a function of the batches in flight (streams per device) and the
packing threads; the numbers behind the default of five streams.   python tools/e2e_batch.py [pairs]"""
import sys, os, time, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import agatha_b200 as ag
n = int(sys.argv[1]) if len(sys.argv) > 1 else 200000
d = ag.synth_pairs(2, 5, n)
p = ag.make_params()
for ba in (8192,):
    for st in (4, 6, 8, 12):
      for stg in (8, 4):
          ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, devices=[0], batch_alns=ba, streams_per_device=st, staging_threads=stg)
          ts = []
          for _ in range(3):
              t0 = time.time()
              res, stats = ag.align_job(d["qbuf"], d["qoff"], d["qlen"], d["tbuf"], d["toff"], d["tlen"], p, devices=[0], batch_alns=ba, streams_per_device=st, staging_threads=stg)
              ts.append(time.time() - t0)
          print("batch", ba, "streams", st, "pack threads", stg, "best %.1f ms  (%.0f al/s)  job-internal %.1f ms, kernel sum %.1f ms, batches %d" % (min(ts) * 1e3, n / min(ts), stats["seconds_total"] * 1e3, stats["seconds_kernel_max"] * 1e3, stats["n_batches"]), flush=True)
