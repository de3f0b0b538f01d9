"""Point-cloud loader for the PointNetVLAD-format ``.bin`` submaps (raw float64
xyz triples), same class name / ``__call__`` contract as the reference's
datasets/pointnetvlad/pnv_raw.py:8-22."""
import numpy as np


class PNVPointCloudLoader:
    def __call__(self, file_pathname: str) -> np.ndarray:
        return self.read_pc(file_pathname)

    def read_pc(self, file_pathname: str) -> np.ndarray:
        pc = np.fromfile(file_pathname, dtype=np.float64).astype(np.float32)
        return pc.reshape(-1, 3)
