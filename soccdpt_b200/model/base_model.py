"""Mirror of the reference's BaseModel (SOccDPT/model/base_model.py:4-46): checkpoint loading with
``strict=False`` (prints the incompatible keys, never raises on a key mismatch) and ``get_device``."""
import torch


class BaseModel(torch.nn.Module):
    def load_net(self, path):
        """Loads a state_dict; unwraps training checkpoints of the form {"optimizer":..., "model":...}."""
        if path is None or not path:
            return
        state = torch.load(path, map_location=torch.device("cpu"))
        if "optimizer" in state:
            print("Loading optimizer state dict")
            state = state["model"]
        incompatible_keys = self.load_state_dict(state, strict=False)
        print("incompatible_keys", incompatible_keys)

    def get_device(self):
        try:
            return next(self.parameters()).device
        except Exception as ex:  # parameter-less module (the bare SOccDPT base class)
            print("No device found, using CPU", ex)
            return torch.device("cpu")
