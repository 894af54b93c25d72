import os
here = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(here, 'e5_bucket_rank.py')).read())
exec(open(os.path.join(here, 'e6_latency.py')).read())
