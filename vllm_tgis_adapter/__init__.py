"""Drop-in package name of the reference (`python -m vllm_tgis_adapter`); everything lives in vllm_tgis_adapter_b200."""
__version__ = "0.1.0+b200"
