"""Drop-in for the reference's ``acquisition`` module: put this directory first on ``sys.path`` and
``Settings.postProcessing`` (reference ``initialize.py:456``, ``:484``) picks up the B200 path."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from softgnss_python_b200.acquisition import (AcquisitionResult, acquisition, preRun,  # noqa: E402,F401
                                              showChannelStatus)
