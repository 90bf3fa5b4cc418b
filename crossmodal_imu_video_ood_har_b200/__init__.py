"""Import alias for the package in ``crossmodal-imu-video-ood-har_b200/`` (a hyphen is not valid in
a Python module name).  ``import crossmodal_imu_video_ood_har_b200 as cm`` executes that package's
``__init__`` under this name; submodules resolve inside the hyphenated directory."""
import os as _os

_real = _os.path.join(_os.path.dirname(_os.path.dirname(_os.path.abspath(__file__))),
                      "crossmodal-imu-video-ood-har_b200")
__path__ = [_real]
with open(_os.path.join(_real, "__init__.py")) as _f:
    exec(compile(_f.read(), _os.path.join(_real, "__init__.py"), "exec"))
