"""Command-line entry (reference source/main.py:5-21):  python3 main.py <setup_module>

Run from ``source/`` with ``../setups`` on the path, exactly as the reference.  Under
``torchrun`` (one process per GPU) the process group replaces ``MPI.COMM_WORLD``.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, '../setups')
sys.path.insert(0, os.path.join(os.path.dirname(_HERE), 'setups'))
sys.path.insert(0, os.path.dirname(_HERE))
sys.path.insert(0, _HERE)


def _init_process_group():
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        import torch
        import torch.distributed as dist
        local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))


def main(argv):
    if len(argv) < 2:
        raise SystemExit("usage: python3 main.py <setup_module>")
    _init_process_group()
    from shakti_b200.fem import comm_world
    comm = comm_world()

    # import the setup module named on the command line
    setup = importlib.import_module(argv[1])

    # initialise the model object with the communicator
    md = setup.initialize(comm)

    # solve; results are written to md.results_name
    md.solve()


if __name__ == "__main__":
    main(sys.argv)
