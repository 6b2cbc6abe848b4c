"""Config validation helper with the behaviour of the reference's check_params
(openchem/utils/utils.py:108-136): required keys must exist and have the declared type, optional keys are
type-checked when present, unknown keys are tolerated."""
from __future__ import annotations


def check_params(config, required_dict, optional_dict):
    if required_dict is None or optional_dict is None:
        return
    for spec, must_exist in ((required_dict, True), (optional_dict, False)):
        for key, kind in spec.items():
            if key not in config:
                if must_exist:
                    raise ValueError("{} parameter has to be specified".format(key))
                continue
            value = config[key]
            if isinstance(kind, list):
                if kind and value not in kind:
                    raise ValueError("{} has to be one of {}".format(key, kind))
            elif kind is not None and not isinstance(value, kind):
                raise ValueError("{} has to be of type {}".format(key, kind))
