"""posematching/procrustes.py of the reference, same signatures: procrustes (:150-259), superimpose (:7-54),
superimpose_old (:68-106).  The superimposition runs in hpusm_procrustes_batch."""
import numpy as np

from .._dev import engine, up, down


def procrustes(X, Y, scaling=True, reflection='best'):
    X = np.asarray(X, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    d, Z, tf = engine.procrustes_batch(up(X[None], np.float64), up(Y[None], np.float64), scaling, reflection)
    tf = down(tf)[0]
    tform = {'rotation': tf[:4].reshape(2, 2), 'scale': (tf[4] if scaling else 1), 'translation': tf[5:7]}
    return float(down(d)[0]), down(Z)[0], tform


def superimpose(input, model, plot=False, input_image=None, model_image=None):
    input = np.asarray(input)
    model = np.asarray(model)
    keep = [k for k in range(0, 18) if not (((model[k][0] == 0) and (model[k][1] == 0)) or
                                            ((input[k][0] == 0) and (input[k][1] == 0)))]
    model = model[keep]
    input = input[keep]
    (d, Z, m) = procrustes(input, model, False)
    Z[:, 1] = Z[:, 1] + (min(model[:, 1]) - min(Z[:, 1]))
    return (Z, model)


def superimpose_old(input, model, plot=False, input_image=None, model_image=None):
    (d, Z, m) = procrustes(input, model, False)
    voet_index = 10 if input[10][1] >= input[13][1] else 13
    Z[:, 1] = Z[:, 1] + (input[voet_index][1] - Z[voet_index][1])
    return Z, model
