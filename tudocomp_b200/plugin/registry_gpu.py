###############################################################################
# Registry configuration for the `tdc` binaries built by build_tdc.sh, in the format of the reference's
# etc/registry_config.py (evaluated by the reference's own etc/genregistry.py).
#
# It is the reference's registry trimmed to the compressors that build offline (no SDSL / Boost / Judy users),
# plus ONE added line per GPU plugin: the GPU text index is registered as an additional `textds` choice, exactly
# where etc/registry_config.py:73-75 lists TextDS.
#
#   TDC_GPU_MODE == "mixed"   reference TextDS (default) + GpuTextDS:   -a "lzss_lcp(huff, gpu)", -a "bwt(gpu)"
#   TDC_GPU_MODE == "only"    GpuTextDS only and the default:           -a "lzss_lcp(coder=huff)" runs on the GPU and
#                                                                       the archive header equals the reference's
#   TDC_GPU_MODE == "none"    the unmodified reference subset (tdc_ref, the byte-identity checker)
###############################################################################
import os

mode = os.environ.get("TDC_GPU_MODE", "mixed")

coders = [
    AlgorithmConfig(name="ASCIICoder", header="coders/ASCIICoder.hpp"),
    AlgorithmConfig(name="BitCoder", header="coders/BitCoder.hpp"),
    AlgorithmConfig(name="HuffmanCoder", header="coders/HuffmanCoder.hpp"),
]

sa = [AlgorithmConfig(name="SADivSufSort", header="ds/SADivSufSort.hpp")]
phi = [AlgorithmConfig(name="PhiFromSA", header="ds/PhiFromSA.hpp")]
plcp = [AlgorithmConfig(name="PLCPFromPhi", header="ds/PLCPFromPhi.hpp")]
lcp = [AlgorithmConfig(name="LCPFromPLCP", header="ds/LCPFromPLCP.hpp")]
isa = [AlgorithmConfig(name="ISAFromSA", header="ds/ISAFromSA.hpp")]

cpu_textds = [AlgorithmConfig(name="TextDS", header="ds/TextDS.hpp", sub=[sa, phi, plcp, lcp, isa])]
# the added line: a GPU-backed text index (tudocomp_b200/plugin/include/tudocomp_gpu/GpuTextDS.hpp)
gpu_textds = [AlgorithmConfig(name="GpuTextDS", header="../tudocomp_gpu/GpuTextDS.hpp")]

textds = {"mixed": cpu_textds + gpu_textds, "only": gpu_textds, "none": cpu_textds}[mode]

tdc.compressors = [
    AlgorithmConfig(name="RunLengthEncoder", header="compressors/RunLengthEncoder.hpp"),
    AlgorithmConfig(name="LiteralEncoder", header="compressors/LiteralEncoder.hpp", sub=[coders]),
    AlgorithmConfig(name="LZSSLCPCompressor", header="compressors/LZSSLCPCompressor.hpp", sub=[coders, textds]),
    AlgorithmConfig(name="MTFCompressor", header="compressors/MTFCompressor.hpp"),
    AlgorithmConfig(name="NoopCompressor", header="compressors/NoopCompressor.hpp"),
    AlgorithmConfig(name="BWTCompressor", header="compressors/BWTCompressor.hpp", sub=[textds]),
    AlgorithmConfig(name="ChainCompressor", header="../tudocomp_driver/ChainCompressor.hpp"),
]

tdc.generators = [
    AlgorithmConfig(name="FibonacciGenerator", header="generators/FibonacciGenerator.hpp"),
    AlgorithmConfig(name="ThueMorseGenerator", header="generators/ThueMorseGenerator.hpp"),
    AlgorithmConfig(name="RandomUniformGenerator", header="generators/RandomUniformGenerator.hpp"),
    AlgorithmConfig(name="RunRichGenerator", header="generators/RunRichGenerator.hpp"),
]
