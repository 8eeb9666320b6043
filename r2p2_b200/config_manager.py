"""ConfigManager with the reference's behaviour (r2p2/config_manager/config_manager.py:3-60).

Scenario JSON -> dict; CLI-style ``params`` override scenario keys; the controller JSON(s) named by the
scenario's ``controller`` key (a path or a list of paths) are loaded into ``config['controllers']``;
scenario keys then override same-named controller keys.  Unlike the reference, ``params=None`` is accepted
(the reference crashes on it) and relative paths are resolved against the scenario file's
directory when they do not exist relative to the cwd, so the tool no longer has to be run from ``r2p2/``.
"""
import json
import os


class ConfigManager:
    def __init__(self, scenario_config="../conf/scenario-default.json", controller_config=None, params=None):
        self.base = os.path.dirname(os.path.abspath(scenario_config))
        self.config = self._get_json(scenario_config)
        self.override_with_cli_params(params or {})
        if controller_config:
            self.config["controllers"] = [self._get_json(controller_config)]
        else:
            self._load_controller_conf()
        self._override_with_scen()

    def override_with_cli_params(self, params):
        self.config = {**self.config, **params}

    def get_config(self):
        return self.config

    def resolve(self, path):
        """Path as written in the JSON (relative to cwd, like the reference) or next to the scenario file."""
        if os.path.exists(path):
            return path
        for cand in (os.path.join(self.base, path), os.path.join(self.base, os.path.basename(path)),
                     os.path.join(self.base, "..", path.replace("../", "", 1))):
            if os.path.exists(cand):
                return os.path.normpath(cand)
        return path

    def _get_json(self, path):
        with open(self.resolve(path) if hasattr(self, "base") else path, "r") as f:
            return json.load(f)

    def _load_controller_conf(self):
        if "controller" in self.config:
            paths = self.config["controller"]
            if not isinstance(paths, list):
                paths = [paths]
            for path in paths:
                self.config.setdefault("controllers", []).append(self._get_json(path))

    def _override_with_scen(self):
        for controller in self.config.get("controllers", []):
            for key in controller:
                if key in self.config:
                    controller[key] = self.config[key]
