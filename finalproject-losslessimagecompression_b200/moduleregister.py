"""Name -> class lookup used by the YAML configs (reference: moduleregister.py:1-22).

The reference keeps ONE class-level dict shared by every registry subclass, so names are global
(`Register.get('IDFlows')` works from any subclass).  The same behaviour is kept because the
configs rely on it (e.g. trainer.py:204 `NNFlows.get(model.pop('name'))`).
"""


class Register:
    record: dict = {}

    @classmethod
    def register(cls, obj):
        Register.record[obj.__name__] = obj
        return obj

    @classmethod
    def get(cls, key):
        try:
            return Register.record[key]
        except KeyError:
            raise Exception(f"Can not find object {key}") from None
