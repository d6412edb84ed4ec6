class Data:
    def __init__(self, x=None, edge_index=None, y=None):
        self.x, self.edge_index, self.y = x, edge_index, y


class DataLoader:          # train*.py:6 imports the name
    pass
