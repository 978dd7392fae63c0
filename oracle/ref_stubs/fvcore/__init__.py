__version__ = "0.1.5.stub"
