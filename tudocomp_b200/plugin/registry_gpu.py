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
# added lines (mixed registry): GPU-backed PROVIDERS next to the reference's, selectable one by one inside the unchanged
# TextDS — `textds(sa=gpu)`, `textds(sa=gpu, lcp=gpu, isa=gpu)` (tudocomp_gpu/GpuProviders.hpp; the provider concept of
# ds/TextDS.hpp:23-29).  Phi / PLCP have GPU classes too (GpuPhi, GpuPLCP); they are left out of THIS registry only to keep
# the number of TextDS<...> instantiations of the offline build at 8 instead of 32.
if mode == "mixed":
    sa_m = sa + [AlgorithmConfig(name="GpuSA", header="../tudocomp_gpu/GpuProviders.hpp")]
    lcp_m = lcp + [AlgorithmConfig(name="GpuLCP", header="../tudocomp_gpu/GpuProviders.hpp")]
    isa_m = isa + [AlgorithmConfig(name="GpuISA", header="../tudocomp_gpu/GpuProviders.hpp")]
    cpu_textds = [AlgorithmConfig(name="TextDS", header="ds/TextDS.hpp", sub=[sa_m, phi, plcp, lcp_m, isa_m])]
# the added line: a GPU-backed text index (tudocomp_b200/plugin/include/tudocomp_gpu/GpuTextDS.hpp)
gpu_textds = [AlgorithmConfig(name="GpuTextDS", header="../tudocomp_gpu/GpuTextDS.hpp")]

textds = {"mixed": cpu_textds + gpu_textds, "only": gpu_textds, "none": cpu_textds}[mode]

# lcpcomp (etc/registry_config.py:135-169) with the strategies / decoders that build offline: BoostHeap and
# PLCPStrategy need Boost.  The reference only allows an uncompressed (writable) LCP provider here; GpuArray is a
# DynamicIntVector at width 32 unless `compress` narrows it, and the strategies overwrite entries in place either way.
lcpcomp_coders = [
    AlgorithmConfig(name="ASCIICoder", header="coders/ASCIICoder.hpp"),
    AlgorithmConfig(name="HuffmanCoder", header="coders/HuffmanCoder.hpp"),
]  # (SLECoder left out only to keep the offline build short)
lcpcomp_comp = [
    AlgorithmConfig(name="lcpcomp::MaxHeapStrategy", header="compressors/lcpcomp/compress/MaxHeapStrategy.hpp"),
    AlgorithmConfig(name="lcpcomp::MaxLCPStrategy", header="compressors/lcpcomp/compress/MaxLCPStrategy.hpp"),
    AlgorithmConfig(name="lcpcomp::ArraysComp", header="compressors/lcpcomp/compress/ArraysComp.hpp"),
    AlgorithmConfig(name="lcpcomp::PLCPPeaksStrategy", header="compressors/lcpcomp/compress/PLCPPeaksStrategy.hpp"),
]
lcpcomp_dec = [
    AlgorithmConfig(name="lcpcomp::ScanDec", header="compressors/lcpcomp/decompress/ScanDec.hpp"),
    AlgorithmConfig(name="lcpcomp::CompactDec", header="compressors/lcpcomp/decompress/CompactDec.hpp"),
]  # (DecodeForwardQueueListBuffer / MultimapBuffer left out only to keep the offline build short: 24 instantiations per textds)
lcp_uncompressed = [AlgorithmConfig(name="LCPFromPLCP", header="ds/LCPFromPLCP.hpp")]
lcpcomp_cpu_textds = [AlgorithmConfig(name="TextDS", header="ds/TextDS.hpp", sub=[sa, phi, plcp, lcp_uncompressed, isa])]
lcpcomp_textds = {"mixed": lcpcomp_cpu_textds + gpu_textds, "only": gpu_textds, "none": lcpcomp_cpu_textds}[mode]
# LCPCompressor::meta() hard-codes TextDS<> as the default of its `textds` option (compressors/LCPCompressor.hpp:91), so a
# registry without the CPU TextDS cannot hold it: lcpcomp is registered in the reference subset and in the mixed registry
# (select the GPU index with `lcpcomp(..., textds=gpu)`), not in the GPU-only one.
lcpcomp = [] if mode == "only" else [
    AlgorithmConfig(name="LCPCompressor", header="compressors/LCPCompressor.hpp", sub=[lcpcomp_coders, lcpcomp_comp, lcpcomp_dec, lcpcomp_textds]),
]

# The stream stages behind the BWT (config 3, `bwt:mtf:rle:encode(huff)`): in the GPU-only registry the same names resolve
# to the GPU classes of tudocomp_gpu/GpuStreamStages.hpp (same Meta, same bytes); elsewhere to the reference's.
if mode == "only":
    stream_stages = [
        AlgorithmConfig(name="GpuRunLengthEncoder", header="../tudocomp_gpu/GpuStreamStages.hpp"),
        AlgorithmConfig(name="GpuLiteralEncoder", header="../tudocomp_gpu/GpuStreamStages.hpp", sub=[coders]),
        AlgorithmConfig(name="GpuMTFCompressor", header="../tudocomp_gpu/GpuStreamStages.hpp"),
    ]
else:
    stream_stages = [
        AlgorithmConfig(name="RunLengthEncoder", header="compressors/RunLengthEncoder.hpp"),
        AlgorithmConfig(name="LiteralEncoder", header="compressors/LiteralEncoder.hpp", sub=[coders]),
        AlgorithmConfig(name="MTFCompressor", header="compressors/MTFCompressor.hpp"),
    ]

tdc.compressors = lcpcomp + stream_stages + [
    AlgorithmConfig(name="LZSSLCPCompressor", header="compressors/LZSSLCPCompressor.hpp", sub=[coders, textds]),
    AlgorithmConfig(name="NoopCompressor", header="compressors/NoopCompressor.hpp"),
    AlgorithmConfig(name="BWTCompressor", header="compressors/BWTCompressor.hpp", sub=[textds]),
    # `a:b` = chain(a, b): the GPU-only registry keeps the bytes between two GPU stages in device memory
    (AlgorithmConfig(name="GpuChainCompressor", header="../tudocomp_gpu/GpuChainCompressor.hpp") if mode == "only" else
     AlgorithmConfig(name="ChainCompressor", header="../tudocomp_driver/ChainCompressor.hpp")),
]

tdc.generators = [
    AlgorithmConfig(name="FibonacciGenerator", header="generators/FibonacciGenerator.hpp"),
    AlgorithmConfig(name="ThueMorseGenerator", header="generators/ThueMorseGenerator.hpp"),
    AlgorithmConfig(name="RandomUniformGenerator", header="generators/RandomUniformGenerator.hpp"),
    AlgorithmConfig(name="RunRichGenerator", header="generators/RunRichGenerator.hpp"),
]
