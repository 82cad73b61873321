from vllm_tgis_adapter_b200.__main__ import main

if __name__ == "__main__":
    main()
