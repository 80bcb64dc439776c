"""placeholder"""
class EngineError(RuntimeError):
    pass
class Engine:
    pass
def library_path():
    return None
