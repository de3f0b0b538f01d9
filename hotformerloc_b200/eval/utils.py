"""Evaluation pickle names per dataset (reference: eval/utils.py:1-36)."""

_SPLITS = {
    'Oxford': (['oxford', 'university', 'residential', 'business'], '{}_evaluation_database.pickle',
               '{}_evaluation_query.pickle'),
    'MulRan': (['DCC', 'Sejong'], '{}_database.pickle', '{}_queries.pickle'),
    'CSWildPlaces': (['Karawatha', 'Venman', 'QCAT', 'Samford'],
                     'CSWildPlaces_{}_evaluation_database.pickle',
                     'CSWildPlaces_{}_evaluation_query.pickle'),
    'WildPlaces': (['Karawatha', 'Venman'], '{}_evaluation_database.pickle',
                   '{}_evaluation_query.pickle'),
}


def get_query_database_splits(params):
    name = params.dataset_name
    if name == 'CSCampus3D':
        return ['umd_evaluation_database.pickle'], ['umd_evaluation_query_v2.pickle']
    key = name if name in _SPLITS else ('CSWildPlaces' if 'CSWildPlaces' in name else
                                        ('WildPlaces' if 'WildPlaces' in name else None))
    if key is None:
        raise NotImplementedError(f'Dataset {name} has no splits implemented')
    locs, db, q = _SPLITS[key]
    return [db.format(l) for l in locs], [q.format(l) for l in locs]
