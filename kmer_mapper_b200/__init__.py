"""kmer_mapper_b200 -- B200-native drop-in for kmer_mapper's read -> k-mer -> index lookup -> count path.

Module layout mirrors the reference package so that ``from kmer_mapper.X import Y`` becomes
``from kmer_mapper_b200.X import Y``:

  mapper                  map_kmers_to_graph_index, in_graph_index, in_graph_index_no_memory_maps (mapper.pyx)
  util                    get_kmer_hashes_from_chunk_sequence, _get_kmer_index_from_args (util.py)
  gpu_counter             GpuCounter (gpu_counter.py)
  encodings               BaseEncoding, ACTGTwoBitEncoding, SimpleEncoding, twobit_swap (encodings.py)
  command_line_interface  main, run_argument_parser, map_bnp, map_cpu, map_gpu (command_line_interface.py)

plus what the reference gets from third-party packages on this path: ``kmer_index.KmerIndex`` (the
index data model and .npz format), ``reader`` (chunked FASTA/FASTQ(.gz) -> flat bases + offsets),
``device`` (handles over the C ABI) and ``distributed`` (read sharding + count all-reduce).
All arithmetic runs in hand-written sm_100a CUDA kernels behind include/kmer_mapper_b200.h.
"""
__version__ = "0.1.0"
