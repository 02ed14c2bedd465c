#!/usr/bin/env python
"""tapir_compute.py -- same command line as the reference's bin/tapir_compute.py; the PI arithmetic runs
in libtapir_cuda (sm_100a) over all loci at once.  Example (README.rst:131-137 of the reference):

    tapir_compute.py Input_Folder/ Input.tree --output Output_Directory \\
        --intervals=32-42,88-98,95-105,164-174 --times=37,93,100,170 --site-rates
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from tapir_b200.driver import main  # noqa: E402

if __name__ == '__main__':
    sys.exit(main())
