"""Per-split evaluation report -- the entry points of the reference's
``eval/pnv_evaluate_splits.py`` (evaluate :32, evaluate_dataset :81, print_eval_stats :335,
pnv_write_eval_stats :347, CLI :373-431).

Same embedding hot path as :mod:`pnv_evaluate` (batched device octree build, B200 forward,
sharded exact top-k); only the bookkeeping differs: every (database run, query run) pair is
reported under the name of its split (the grand-parent directory of the run's first submap
file), and a location's ``'average'`` entry exists only when it has more than one pair.
"""
from __future__ import annotations

import os
import pickle

import numpy as np

from ..misc.utils import TrainingParams
from . import pnv_evaluate as _pe
from .utils import get_query_database_splits

get_latent_vectors = _pe.get_latent_vectors
get_recall = _pe.get_recall


def _split_of(run: dict) -> str:
    """'<split>/<run dir>/<file>' -> '<split>' (reference :117-119)."""
    return os.path.split(os.path.split(run[0]['query'])[0])[0]


def evaluate_dataset(model, device, params: TrainingParams, database_sets, query_sets,
                     log: bool = False, model_name: str = 'model', show_progress: bool = False):
    model.eval()
    db_vecs = [get_latent_vectors(model, s, device, params) for s in database_sets]
    q_vecs = [get_latent_vectors(model, s, device, params) for s in query_sets]
    campus = 'CSCampus3D' in params.dataset_name
    stats, recalls, oprs, mrrs = {}, [], [], []
    for i in range(len(database_sets)):
        if campus and i != 1:                       # CSCampus3D reports the aerial-only database
            continue
        for j in range(len(query_sets)):
            if (i == j and params.skip_same_run) or db_vecs[i] is None or q_vecs[j] is None:
                continue
            name = _split_of(database_sets[i]) + f'_idx{i}' if campus else _split_of(query_sets[j])
            r, opr, mrr = get_recall(i, j, db_vecs, q_vecs, query_sets, database_sets, log=log,
                                     model_name=model_name)
            recalls.append(np.array(r))
            oprs.append(opr)
            mrrs.append(mrr)
            stats[name] = {'ave_one_percent_recall': opr, 'ave_recall': r, 'ave_mrr': mrr}
    if len(recalls) > 1:
        stats['average'] = {'ave_one_percent_recall': np.mean(oprs),
                            'ave_recall': np.sum(recalls, axis=0) / len(recalls),
                            'ave_mrr': np.mean(mrrs)}
    return stats


def evaluate(model, device, params: TrainingParams, log: bool = False, model_name: str = 'model',
             show_progress: bool = False):
    db_files, q_files = get_query_database_splits(params)
    assert len(db_files) == len(q_files)
    pos = 1 if 'CSWildPlaces' in params.dataset_name else 0
    stats, per_loc = {}, []
    for db_file, q_file in zip(db_files, q_files):
        loc = db_file.split('_')[pos]
        assert loc == q_file.split('_')[pos], \
            f'Database location: {db_file} does not match query location: {q_file}'
        with open(os.path.join(params.dataset_folder, db_file), 'rb') as f:
            database_sets = pickle.load(f)
        with open(os.path.join(params.dataset_folder, q_file), 'rb') as f:
            query_sets = pickle.load(f)
        s = evaluate_dataset(model, device, params, database_sets, query_sets, log=log,
                             model_name=model_name, show_progress=show_progress)
        stats[loc] = s
        per_loc.append(s['average'] if 'average' in s else s[next(iter(s))])
    stats['average'] = {'average': {
        'ave_one_percent_recall': np.mean([a['ave_one_percent_recall'] for a in per_loc]),
        'ave_recall': np.mean([a['ave_recall'] for a in per_loc], axis=0),
        'ave_mrr': np.mean([a['ave_mrr'] for a in per_loc])}}
    return stats


def _fmt(a) -> str:
    return '    ' + str(a).replace('\n', '\n    ')


def print_eval_stats(stats):
    for ds, splits in stats.items():
        print('Dataset: {}'.format(ds))
        for split, v in splits.items():
            print('    Split: {}'.format(split))
            print('    Avg. top 1% recall: {:.2f}   Avg. MRR: {:.2f}   Avg. recall @N:'.format(
                v['ave_one_percent_recall'], v['ave_mrr']))
            print(_fmt(v['ave_recall']))


def pnv_write_eval_stats(file_name, prefix, stats):
    out = [prefix]
    for ds, splits in stats.items():
        out.append(f'\n[{ds}]\n')
        for split, v in splits.items():
            out.append(f'    Split: [{split}]\n')
            out.append('    AR@1%: {:0.2f}, AR@1: {:0.2f}, MRR: {:0.2f}, AR@N:\n'.format(
                v['ave_one_percent_recall'], v['ave_recall'][0], v['ave_mrr']))
            out.append(_fmt(v['ave_recall']) + '\n')
    out.append('\n------------------------------------------------------------------------\n\n')
    with open(file_name, 'a') as f:
        f.write(''.join(out))


def main(argv=None):
    _pe.main(argv, evaluate_fn=evaluate, print_fn=print_eval_stats, write_fn=pnv_write_eval_stats,
             results_suffix='split_results')


if __name__ == '__main__':
    main()
