"""recorder.Recorder for the b200 engine (reference: recorder.py:4-45): appends
grid size, step count, wall time and process memory to record/record_<date>.txt."""
import datetime
import os

import psutil


class Recorder:

    def __init__(self, space, start_time, savedir):
        if space.MPIrank != 0:
            return
        self.savedir = savedir
        self.space = space
        finished_time = datetime.datetime.now()
        os.makedirs(self.savedir + "record/", exist_ok=True)
        record_path = self.savedir + "record/record_%s.txt" % (datetime.date.today())
        if not os.path.exists(record_path):
            with open(record_path, 'a') as f:
                f.write("{:4}\t{:4}\t{:4}\t{:4}\t{:4}\t\t{:4}\t\t{:4}\t\t{:8}\t{:4}\t\t\t\t{:6}\t{:12}\t{:12}\n\n"
                        .format("Node", "Nx", "Ny", "Nz", "dx", "dy", "dz", "tsteps", "Time", "Method",
                                "VM/Node(GB)", "RM/Node(GB)"))
        me = psutil.Process(os.getpid())
        rss = float(me.memory_info().rss) / 1024 / 1024 / 1024
        vms = float(me.memory_info().vms) / 1024 / 1024 / 1024
        cal_time = finished_time - start_time
        with open(record_path, 'a') as f:
            f.write("{:2d}\t\t{:04d}\t{:04d}\t{:04d}\t{:5.2e}\t{:5.2e}\t{:5.2e}\t{:06d}\t\t{}\t\t{:>6}\t\t{:06.3f}\t\t\t{:06.3f}\n"
                    .format(space.MPIsize, space.Nx, space.Ny, space.Nz, space.dx, space.dy, space.dz,
                            space.tsteps, cal_time, space.method, vms, rss))
        print("Simulation specifications are recorded. {}".format(datetime.datetime.now()))
