class DictConfig(dict): pass
class ListConfig(list): pass
class SCMode: pass
class OmegaConf: pass
