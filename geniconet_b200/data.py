"""Synthetic level-s meshes in the reference's on-disk tensor contract (data.py:64-69,
generate.py:200-203): input [3,5n,2n] = xyz of the P grid vertices, target [9,P+2] = xyz | normals |
Laplacian.  SURVEY.md 8d recipe: an icosphere radially displaced by a smooth seeded field, |xyz| < 1.

Targets are made on the host with numpy (the generate.py:20-43 normals recipe and the uniform
Laplacian) -- dataset preparation, not the hot path.
"""
import numpy as np
import torch

from .ico_geometry import get_icosahedral_grid

_cache = {}


def _topology(s):
    if s not in _cache:
        v, f = get_icosahedral_grid(s)
        V = v.shape[0]
        e = np.concatenate([f[:, [0, 1]], f[:, [1, 2]], f[:, [2, 0]]])
        e = np.concatenate([e, e[:, ::-1]])
        e = np.unique(e, axis=0)
        deg = np.bincount(e[:, 0], minlength=V).astype(np.float64)
        _cache[s] = (v.astype(np.float64), f, e, deg)
    return _cache[s]


def vertex_normals(v, f, eps=1e-10, reference_semantics=False):
    """Area-weighted vertex normals (the generate.py:20-43 recipe, eps clip).

    reference_semantics=True reproduces what generate.py:36-38 actually computes: `v_normals[faces[:, c], :] += f_normals` is a
    buffered fancy-index update, so a vertex that is corner `c` of several faces receives only the LAST of them -- each vertex
    normal is the sum of at most three face normals (one per corner slot), not of all its faces.  The published dataset's normal
    targets were made that way (generate.py:194); on a smooth level-3 mesh the two differ by 2.4 degrees on average.  The default
    is the true accumulation (what the loss's compute_vertex_normals restatement and the CUDA kernel compute)."""
    fn = np.cross(v[f[:, 1]] - v[f[:, 0]], v[f[:, 2]] - v[f[:, 0]])
    vn = np.zeros_like(v)
    for c in range(3):
        if reference_semantics:
            last = np.zeros_like(v)
            last[f[:, c]] = fn                       # plain assignment: the last face per vertex wins, as in the reference
            vn += last
        else:
            np.add.at(vn, f[:, c], fn)
    return vn / np.clip(np.sqrt((vn ** 2).sum(1)), eps, None)[:, None]


def laplacian(v, e, deg):
    acc = np.zeros_like(v)
    np.add.at(acc, e[:, 0], v[e[:, 1]])
    return acc / deg[:, None] - v


def synthetic_mesh(s, sample_idx):
    """(input float32 [3,5n,2n], target float32 [9,P+2]) for one seeded sample."""
    base, f, e, deg = _topology(s)
    g = torch.Generator().manual_seed(1234 + int(sample_idx))
    a = torch.rand(4, generator=g).numpy() * 2 - 1
    w = torch.randn(4, 3, generator=g).numpy() * 2.0
    ph = torch.rand(4, generator=g).numpy() * 2 * np.pi
    field = sum(a[j] * np.sin(base @ w[j] + ph[j]) for j in range(4))
    r = 0.5 * (1.0 + 0.25 * field)
    v = base * r[:, None]
    tgt = np.concatenate([v, vertex_normals(v, f), laplacian(v, e, deg)], axis=1).T.astype(np.float32)   # [9,P+2]
    n = 2 ** s
    x = np.ascontiguousarray(tgt[:3, :-2].reshape(3, 5 * n, 2 * n))
    return torch.from_numpy(x), torch.from_numpy(np.ascontiguousarray(tgt))


def synthetic_batch(s, first_idx, batch):
    xs, ts = zip(*(synthetic_mesh(s, first_idx + i) for i in range(batch)))
    return torch.stack(xs), torch.stack(ts)


# ----------------------------------------------------------------------------------------------------------------------
# The reference's on-disk data path: file listing, `.npz` reader and Dataset classes with the
# reference's names and return conventions (data.py:7-160), plus a pinned-memory device prefetcher for the training loop.
# Host-side code: nothing here runs on the oracle, and the tensors a Dataset yields are CPU tensors exactly like the
# reference's -- `DevicePrefetcher` is what moves them.
# ----------------------------------------------------------------------------------------------------------------------
import os
import re


def natural_key(name):
    """Sort key with the ordering natsort.natsorted gives file names (data.py:20,34): digit runs compare as integers."""
    return [int(tok) if tok.isdigit() else tok for tok in re.split(r'(\d+)', name)]


def _listdir_ext(path, ext):
    return [f for f in sorted(os.listdir(path), key=natural_key) if f.endswith(ext)]


def listFiles(params, data_type, data_instance):
    """data.py:7-38.  dataPthLvl 1: one flat directory (per data_instance for 'enc'/'ftr'); dataPthLvl 2 (ModelNet):
    <dataPth>/<class>/<train|test>/, with 'trn'/'val' mapped to 'train'/'test'."""
    full = []
    lvl = params['ico']['dataPthLvl']
    if lvl == 1:
        if data_type in ('enc', 'ftr'):
            path = os.path.join(params[data_type]['dataPth'], data_instance)
        else:
            path = params[data_type]['dataPth']
        full = [os.path.join(path, f) for f in _listdir_ext(path, params[data_type]['ext'])]
    elif lvl == 2:
        data_instance = {'trn': 'train', 'val': 'test'}.get(data_instance, data_instance)
        for cls in os.listdir(params[data_type]['dataPth']):      # the reference keeps os.listdir order here too
            sub = os.path.join(params[data_type]['dataPth'], cls, data_instance)
            full += [os.path.join(sub, f) for f in _listdir_ext(sub, params[data_type]['ext'])]
    return full


def loadEncFile(params, inFile):
    """data.py:40-46: an encoding saved by save_to_file(..., '.npz') -> tensor of 'arr_0'."""
    ext = os.path.splitext(inFile)[1]
    if ext != '.npz':
        raise ValueError('File format %s not specified for loadEncFile' % ext)
    return torch.tensor(np.load(inFile)['arr_0'])


def loadIcoFile(params, inFile):
    """data.py:48-75.  '.npz': key 'data' = [9, P+2] (xyz | normals | Laplacian, generate.py:200-203) -> (input
    [3,5n,2n] = the xyz rows without the two poles, reshaped with params['ico']['width'] columns; target [9,P+2]).
    '.mat' with a 'variable' image: channels first, /255, target == input.  Anything else: ValueError."""
    ext = params['ico']['ext']
    if ext == '.npz':
        lbl2 = np.load(inFile)['data']
        lbl1 = lbl2[:3, :-2]
        return lbl1.reshape(lbl1.shape[0], -1, params['ico']['width']), lbl2
    if ext == '.mat':
        import scipy.io
        lbl = scipy.io.loadmat(inFile)
        if 'variable' in lbl:
            lbl = np.ascontiguousarray(np.transpose(lbl['variable'], (2, 0, 1))).astype(np.float32)
            lbl[0:3] /= 255.0
            lbl[3:6] = lbl[0:3]
            if np.isnan(lbl).any():                       # the reference's assert can never fire (isnan(lbl.all())); this one can
                raise ValueError('NaN in %s' % inFile)
            return lbl, lbl
        if 'sparse_weights' in lbl:
            raise ValueError('mat file with sparse_weights and sparse_vertices cannot be handled here, use generate.py')
        raise ValueError('content of mat file unhandleable')
    raise ValueError('ico loader for %s not specified' % ext)  # the reference RETURNS this error object (data.py:75); raising is the intent


def write_ico_npz(path, target):
    """The writer side of the contract (generate.py:200-203): np.savez(path, data=[9, P+2])."""
    np.savez(path, data=np.asarray(target, dtype=np.float32))


class createico2icoDataset(torch.utils.data.Dataset):
    """data.py:78-107: every pair is loaded up front; 'train' yields (ico, target), 'test' yields (ico, out path stem, ico)."""

    def __init__(self, params, data_instance):
        self.params = params
        self.icoList = listFiles(params, 'ico', data_instance)
        self.icoPair = [loadIcoFile(params, f) for f in self.icoList]
        if params['process_name'] == 'test':
            self.outIcoPth = os.path.join(params['out']['dataPth'], params[params['model_name']]['data_instance'])
            os.makedirs(self.outIcoPth, exist_ok=True)

    def __getitem__(self, idx):
        ico, outIco = self.icoPair[idx]
        if self.params['process_name'] == 'train':
            return ico, outIco
        if self.params['process_name'] == 'test':
            return ico, os.path.join(self.outIcoPth, os.path.basename(self.icoList[idx]).split('.')[0]), ico
        raise ValueError('%s process on %s model not defined' % (self.params['process_name'], self.params['model_name']))

    def __len__(self):
        return len(self.icoList)


class createico2encDataset(torch.utils.data.Dataset):
    """data.py:109-125: lazy loading; yields (ico, path the encoding is to be written to)."""

    def __init__(self, params, data_instance):
        self.params = params
        self.icoList = listFiles(params, 'ico', data_instance)
        self.encDataPth = os.path.join(params['enc']['dataPth'], data_instance)
        os.makedirs(self.encDataPth, exist_ok=True)

    def __getitem__(self, idx):
        ico, _ = loadIcoFile(self.params, self.icoList[idx])
        stem = os.path.basename(self.icoList[idx]).split('.')[0]
        return ico, os.path.join(self.encDataPth, stem + self.params['enc']['ext'])

    def __len__(self):
        return len(self.icoList)


class createenc2icoDataset(torch.utils.data.Dataset):
    """data.py:127-150: encodings matched to ico files by base name; yields (enc, output path stem, ico)."""

    def __init__(self, params, data_instance):
        self.params = params
        encList = listFiles(params, 'enc', data_instance)
        icoList = listFiles(params, 'ico', data_instance)
        encNames = set(os.path.basename(f) for f in encList)
        self.encList = encList
        self.icoList = [f for f in icoList if os.path.basename(f) in encNames]
        self.outDataPth = os.path.join(params['out']['dataPth'], data_instance)
        os.makedirs(self.outDataPth, exist_ok=True)

    def __getitem__(self, idx):
        enc = loadEncFile(self.params, self.encList[idx])
        icoPath = os.path.join(self.outDataPth, os.path.basename(self.encList[idx]).split('.')[0])
        ico, _ = loadIcoFile(self.params, self.icoList[idx])
        return enc, icoPath, ico

    def __len__(self):
        return len(self.icoList)


class createico2ico_vaeDataset(createico2icoDataset):
    pass


class createico2enc_vaeDataset(createico2encDataset):
    pass


class createenc2ico_vaeDataset(createenc2icoDataset):
    pass


class DevicePrefetcher:
    """Wraps an iterable of (input, target) CPU batches (a DataLoader over the datasets above with pin_memory=True, or any
    generator) and yields device tensors, the next batch's host->device copy running on its own stream while the caller
    computes on the current one.  Two device staging slots; a slot is refilled only after the consumer's stream has passed
    the point where it stopped using it (event recorded at the next __next__).  Works without CUDA only as a pass-through
    for tests (device='cpu')."""

    def __init__(self, batches, device='cuda'):
        self.it = iter(batches)
        self.device = torch.device(device)
        self.cuda = self.device.type == 'cuda'
        if self.cuda:
            self.stream = torch.cuda.Stream(self.device)
            self.slots = [None, None]          # (tensors, ready event)
            self.released = [None, None]       # event: consumer finished with this slot
            self.turn = 0
            self._stage(0)

    def _pin(self, t):
        t = torch.as_tensor(t)
        return t if t.is_pinned() else t.pin_memory()

    def _stage(self, k):
        try:
            batch = next(self.it)
        except StopIteration:
            self.slots[k] = None
            return
        if self.released[k] is not None:
            self.stream.wait_event(self.released[k])
        with torch.cuda.stream(self.stream):
            dev = tuple(self._pin(t).to(self.device, non_blocking=True) if isinstance(t, (torch.Tensor, np.ndarray)) else t
                        for t in batch)
            ev = torch.cuda.Event()
            ev.record(self.stream)
        self.slots[k] = (dev, ev)

    def __iter__(self):
        return self

    def __next__(self):
        if not self.cuda:
            return tuple(torch.as_tensor(t) if isinstance(t, np.ndarray) else t for t in next(self.it))
        k = self.turn
        cur = self.slots[k]
        if cur is None:
            raise StopIteration
        prev = 1 - k
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream(self.device))    # everything queued so far used the previous slot at most
        self.released[prev] = ev
        self._stage(prev)
        dev, ready = cur
        torch.cuda.current_stream(self.device).wait_event(ready)
        for t in dev:
            if isinstance(t, torch.Tensor):
                t.record_stream(torch.cuda.current_stream(self.device))
        self.turn = prev
        return dev
