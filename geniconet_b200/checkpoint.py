"""Checkpoints in the reference's format (run.py:317-409), SURVEY 8f rank 3.

A checkpoint is ``torch.save({'model_state_dict', 'optimizer_state_dict', 'epoch', 'loss', 'misc'})`` at
``<logDir>/savedModel/<modelName>_E<epoch>.pt``; best models carry ``EB<epoch>``.  Loading keeps only the keys the model
has (run.py:359-366), so encoder-only / decoder-only models load from a full autoencoder checkpoint.

The models here keep the reference's attribute names, hence its state-dict keys; the ico layers' own parameter names
(`weight [Cout,Cin,7]`, `bias`) are this build's choice because the real icocnn layout is unpinned -- a
checkpoint written by the original icocnn would need `key_map` / `tensor_map` to translate them.
"""
import glob
import os

import torch

from .data import natural_key


def _epoch_number(epoch):
    """Checkpoint tags are the epoch itself or a one-letter prefix + epoch ('B12' for best models, run.py:326)."""
    return epoch if isinstance(epoch, int) else int(str(epoch)[1:])


def _model_path(params, modelName, epoch):
    return os.path.join(params['logDir'], 'savedModel', modelName + '_E' + str(epoch) + '.pt')


def _best_paths(params, modelName):
    return sorted(glob.glob(os.path.join(params['logDir'], 'savedModel', modelName + '_EB*[0-9]*.pt')), key=natural_key)


def saveModel(params, model, optimizer, epoch, modelName, val_loss, misc):
    """run.py:330-340.  Never overwrites an existing file; returns True when a file was written."""
    path = _model_path(params, modelName, epoch)
    os.makedirs(os.path.dirname(path), exist_ok=True)
    if os.path.exists(path):
        print('%s model with %s epochs, already exists at %s, aborting saving !!' % (modelName, str(epoch), path))
        return False
    torch.save({'model_state_dict': model.state_dict(), 'optimizer_state_dict': optimizer.state_dict(),
                'epoch': _epoch_number(epoch), 'loss': val_loss, 'misc': misc}, path)
    return True


def saveBestModel(params, model, optimizer, epoch, modelName, last_best_loss, last_loss, misc=None):
    """run.py:317-328: on an improved (<=) loss drop all but the five newest best checkpoints, then save 'B<epoch>'."""
    if last_loss[0] <= last_best_loss[0]:
        paths = _best_paths(params, modelName)
        for p in paths[:max(0, len(paths) - 5)]:
            os.remove(p)
        saveModel(params, model, optimizer, 'B' + str(epoch), modelName, last_loss[0], misc)
        last_best_loss[0] = last_loss[0]


def loadModel(params, model, savedEpoch, modelName, optimizer=None, last_best_loss=None, misc=None, key_map=None, tensor_map=None):
    """run.py:342-381.  savedEpoch is a one-element list: [0] selects the newest best checkpoint, otherwise
    '<modelName>_E<savedEpoch[0]>.pt'; it is overwritten with the stored epoch.  Returns False when no file exists.
    Only keys present in the model are loaded.  Unlike the reference, the filtered dict is loaded with strict=False so a
    partial checkpoint really is accepted (run.py:366 would raise on missing keys), and a shape mismatch raises ValueError
    naming the key.  key_map(name) -> name and tensor_map(name, tensor) -> tensor translate foreign checkpoints."""
    if list(savedEpoch) == [0]:
        paths = _best_paths(params, modelName)
        path = paths[-1] if paths else _model_path(params, modelName, 'B*')
    else:
        path = _model_path(params, modelName, savedEpoch[0])
    if not os.path.exists(path):
        print('No saved model exists at %s' % path)
        return False
    checkpoint = torch.load(path, map_location='cpu', weights_only=False)
    model_dict = model.state_dict()
    saved = checkpoint['model_state_dict']
    picked = {}
    for k, v in saved.items():
        k2 = key_map(k) if key_map else k
        if k2 not in model_dict:
            continue
        v2 = tensor_map(k2, v) if tensor_map else v
        if tuple(v2.shape) != tuple(model_dict[k2].shape):
            raise ValueError('checkpoint %s: %s has shape %s, the model expects %s' % (path, k2, tuple(v2.shape), tuple(model_dict[k2].shape)))
        picked[k2] = v2
    model.load_state_dict(picked, strict=False)
    print('Selected %d dict keys out of %d keys' % (len(picked), len(saved)))
    if optimizer is not None and optimizer != []:
        optimizer.load_state_dict(checkpoint['optimizer_state_dict'])
    savedEpoch[0] = checkpoint['epoch']
    if last_best_loss is not None and 'loss' in checkpoint:
        last_best_loss[0] = checkpoint['loss']
    if misc is not None:
        misc.append(checkpoint['misc'])
    for kind in ('out', 'enc', 'ftr'):                    # run.py:377-380
        if kind in params and 'dataPth' in params[kind]:
            params[kind]['dataPth'] = params[kind]['dataPth'].replace('E0', 'EB' + str(savedEpoch[0]))
    return True


def loadMultiModel(params, model, savedEpochs, modelNames):
    """run.py:383-409: fill one model from several checkpoints (e.g. encoder from one run, decoder from another); a key is
    taken from the first checkpoint that has it.  A missing file raises ValueError."""
    remaining = dict(model.state_dict())
    common = {}
    for savedEpoch, modelName in zip(savedEpochs, modelNames):
        path = _model_path(params, modelName, savedEpoch)
        if not os.path.exists(path):
            raise ValueError('No saved model exists at %s' % path)
        saved = torch.load(path, map_location='cpu', weights_only=False)['model_state_dict']
        for k, v in saved.items():
            if k in remaining:
                common[k] = v
                del remaining[k]
    model.load_state_dict(common, strict=False)
    return True
