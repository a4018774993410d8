"""gymutil subset used by legged_gym/utils/helpers.py and base_task.py."""


def parse_device_str(s):
    if ":" in s:
        kind, idx = s.split(":")
        return kind, int(idx)
    return s, 0


def parse_sim_config(cfg, sim_params):
    """Copy the fields the reference reads back (dt, substeps, physx.*) onto SimParams."""
    for k, v in cfg.items():
        if k == "physx":
            for kk, vv in v.items():
                setattr(sim_params.physx, kk, vv)
        elif k == "gravity":
            sim_params.gravity = tuple(v)
        else:
            setattr(sim_params, k, v)


def parse_arguments(*a, **k):
    raise RuntimeError("stub: build the argparse.Namespace yourself (see oracle/ref_runner.py)")


class WireframeSphereGeometry:
    def __init__(self, *a, **k):
        pass


def draw_lines(*a, **k):
    pass
