"""p50 of ll_fa_update_transformation (f-1 + f-3 in one launch) on bench.py's synthetic frame; LLB200_FA_STAGE=0 keeps the
previous sweep's clouds in global memory (A/B of the TMA bulk-copy staging). Also a surf-heavy case: transformCur far off,
so that the surf loop runs many search iterations over the 5.5 k-point cloud."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
r = bench.run_next_rows(with_cpu=False)
print("stage" if os.environ.get("LLB200_FA_STAGE", "1") != "0" else "no-stage", json.dumps(r.get("f1_f3_fa_update_transformation")),
      json.dumps(r.get("f1_f3_fa_update_transformation_far")))
