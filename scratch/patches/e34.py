import os
here = os.path.dirname(os.path.abspath(__file__))
exec(open(os.path.join(here, 'e3_loads8.py')).read())
exec(open(os.path.join(here, 'e4_bbox.py')).read())
