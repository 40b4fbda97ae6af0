"""`Data`: attribute bag with item access, as the reference uses it
(tuned_SIGN.py:117-131, :182-185: ``Data(x=..., y=...)`` then ``data[f'x{k}'] = ...``)."""


class Data:
    def __init__(self, x=None, edge_index=None, **kwargs):
        # construct_pyg_graph passes x and edge_index positionally (reference utils.py:313)
        store = {}
        if x is not None:
            store['x'] = x
        if edge_index is not None:
            store['edge_index'] = edge_index
        store.update(kwargs)
        self.__dict__['_store'] = store

    def __getattr__(self, key):
        store = self.__dict__['_store']
        if key in store:
            return store[key]
        if key in ('x', 'y', 'edge_index', 'edge_weight', 'edge_attr', 'pos'):
            return None
        raise AttributeError(key)

    def __setattr__(self, key, value):
        self.__dict__['_store'][key] = value

    def __getitem__(self, key):
        return self.__dict__['_store'][key]

    def __setitem__(self, key, value):
        self.__dict__['_store'][key] = value

    def __delitem__(self, key):
        del self.__dict__['_store'][key]

    def __contains__(self, key):
        return key in self.__dict__['_store']

    def pop(self, key):
        return self.__dict__['_store'].pop(key)

    def keys(self):
        return list(self.__dict__['_store'].keys())

    @property
    def num_nodes(self):
        st = self.__dict__['_store']
        if 'num_nodes' in st:
            return st['num_nodes']
        return st['x'].shape[0]
