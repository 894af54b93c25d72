"""p50 of ll_fa_update_transformation (f-1 + f-3 in one launch) on bench.py's synthetic frame;
LLB200_FA_STAGE=1 stages the previous sweep's clouds in shared memory by TMA bulk copies (the A/B of profiles/r02_ab_records.txt). Also a surf-heavy case: transformCur far off,
so that the surf loop runs many search iterations over the 5.5 k-point cloud."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
r = bench.run_next_rows(with_cpu=False)
print("staged (cp.async.bulk)" if os.environ.get("LLB200_FA_STAGE", "0") == "1" else "global loads", json.dumps(r.get("f1_f3_fa_update_transformation")),
      json.dumps(r.get("f1_f3_fa_update_transformation_far")))
