mkdir -p gpurun_out
timeout 300 python tools/profile_vae.py > gpurun_out/steps10_vae.log 2>&1
python tools/summarize_steps.py gpurun_out/steps10_vae.log 22; grep "vae decode" gpurun_out/steps10_vae.log
