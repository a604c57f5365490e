class _Catalog(dict):
    def get(self, name, default=None):
        if name not in self:
            self[name] = type("Metadata", (), {"name": name})()
        return self[name]


MetadataCatalog = _Catalog()
DatasetCatalog = _Catalog()
