"""``.pcd`` submap loader (CS-Wild-Places / Wild-Places).  The reference reads
these through open3d (datasets/CSWildPlaces/CSWildPlaces_raw.py:14-23); open3d
is not available offline, so this is a small native PCD v0.7 reader (ascii,
binary and the x/y/z fields only) returning the same (N,3) float32 array."""
import numpy as np

_TYPES = {('F', 4): np.float32, ('F', 8): np.float64, ('I', 4): np.int32, ('I', 2): np.int16,
          ('I', 1): np.int8, ('U', 4): np.uint32, ('U', 2): np.uint16, ('U', 1): np.uint8}


def read_pcd_xyz(path: str) -> np.ndarray:
    with open(path, 'rb') as f:
        header = {}
        while True:
            line = f.readline()
            if not line:
                raise ValueError(f'{path}: truncated PCD header')
            tok = line.decode('ascii', 'replace').strip().split()
            if not tok or tok[0].startswith('#'):
                continue
            header[tok[0].upper()] = tok[1:]
            if tok[0].upper() == 'DATA':
                break
        fields = header['FIELDS']
        sizes = [int(s) for s in header['SIZE']]
        types = header['TYPE']
        counts = [int(c) for c in header.get('COUNT', ['1'] * len(fields))]
        npts = int(header['POINTS'][0])
        mode = header['DATA'][0].lower()
        names, fmts = [], []
        for name, s, t, c in zip(fields, sizes, types, counts):
            for k in range(c):
                names.append(name if c == 1 else f'{name}_{k}')
                fmts.append(_TYPES[(t.upper(), s)])
        if mode == 'ascii':
            data = np.loadtxt(f, dtype=np.float64, ndmin=2)
            cols = [names.index(a) for a in ('x', 'y', 'z')]
            xyz = data[:npts, cols]
        elif mode == 'binary' and names == ['x', 'y', 'z'] and all(ft is np.float32 for ft in fmts):
            # xyz-only float32 records: one reshape (float32 -> float64 -> float32 below is the identity)
            xyz = np.frombuffer(f.read(npts * 12), dtype=np.float32, count=npts * 3).reshape(npts, 3)
            return np.ascontiguousarray(xyz[np.isfinite(xyz).all(1)])
        elif mode == 'binary':
            dt = np.dtype({'names': names, 'formats': fmts})
            rec = np.frombuffer(f.read(npts * dt.itemsize), dtype=dt, count=npts)
            xyz = np.stack([rec['x'], rec['y'], rec['z']], 1)
        else:
            raise NotImplementedError(f'{path}: PCD DATA {mode} is not supported')
    xyz = np.asarray(xyz, dtype=np.float64)
    xyz = xyz[np.isfinite(xyz).all(1)]
    return np.float32(xyz)


class CSWildPlacesPointCloudLoader:
    def __call__(self, file_pathname: str) -> np.ndarray:
        return self.read_pc(file_pathname)

    def read_pc(self, file_pathname: str) -> np.ndarray:
        assert file_pathname.endswith('.pcd')
        return read_pcd_xyz(file_pathname)
