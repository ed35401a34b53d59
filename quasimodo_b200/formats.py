"""File formats at the drop-in boundary: the count TSV that stands where the text pileup
stood ({seq_dir}/pileup/{sample}.{ref}.mpileup family, schema B.3) and the caller VCF
({snpcall_dir}/bcftools/{sample}.{ref}.bcftools.vcf, constraints B.4).  Host-side text output only."""
import numpy as np

BASES = "ACGT"
COUNT_COLUMNS = ["chrom", "pos", "ref", "depth", "A_f", "C_f", "G_f", "T_f", "N_f", "del_f", "A_r", "C_r", "G_r", "T_r",
                 "N_r", "del_r", "ins_start", "del_start", "raw_depth", "read_starts"]


def write_count_tsv(path, genome, rows):
    """rows: int32 [l_pac, 16] (qm_sample_counts_host).  depth = bases passing the BQ filter (channels 0-4, 6-10)"""
    with open(path, "w") as fh:
        fh.write("\t".join(COUNT_COLUMNS) + "\n")
        off = 0
        for name, ln in zip(genome.names, genome.lens):
            blk = rows[off:off + ln]
            depth = blk[:, 0:5].sum(1) + blk[:, 6:11].sum(1)
            ref = genome.codes[off:off + ln]
            for i in range(ln):
                r = blk[i]
                fh.write(f"{name}\t{i + 1}\t{BASES[ref[i]]}\t{depth[i]}\t" + "\t".join(str(int(x)) for x in r) + "\n")
            off += ln


def vcf_header(genome, sample, ref_path="ref.fa"):
    h = ["##fileformat=VCFv4.2", "##FILTER=<ID=PASS,Description=\"All filters passed\">",
         "##source=quasimodo_b200 (threshold caller on bcftools-mpileup-style counts; QUAL is not bcftools call QUAL)",
         f"##reference=file://{ref_path}"]
    h += [f"##contig=<ID={n},length={ln}>" for n, ln in zip(genome.names, genome.lens)]
    h += ["##INFO=<ID=DP,Number=1,Type=Integer,Description=\"Raw read depth\">",
          "##INFO=<ID=AF,Number=1,Type=Float,Description=\"Alternate allele fraction among bases passing the BQ filter\">",
          "##INFO=<ID=DP4,Number=4,Type=Integer,Description=\"ref-forward, ref-reverse, alt-forward, alt-reverse bases\">",
          "##FORMAT=<ID=GT,Number=1,Type=String,Description=\"Genotype\">",
          f"#CHROM\tPOS\tID\tREF\tALT\tQUAL\tFILTER\tINFO\tFORMAT\t{sample}"]
    return h


def write_vcf(path, genome, sample, calls):
    """calls: CALL_DTYPE array sorted by position (qm_call_snps)"""
    with open(path, "w") as fh:
        fh.write("\n".join(vcf_header(genome, sample)) + "\n")
        for c in calls:
            qual = f"{float(c['qual']):.3f}".rstrip("0").rstrip(".")
            info = f"DP={int(c['dp'])};AF={float(c['af']):.3f};DP4={int(c['ad_ref_f'])},{int(c['ad_ref_r'])},{int(c['ad_alt_f'])},{int(c['ad_alt_r'])}"
            fh.write(f"{genome.names[int(c['rid'])]}\t{int(c['pos']) + 1}\t.\t{BASES[int(c['ref'])]}\t{BASES[int(c['alt'])]}\t{qual}\tPASS\t{info}\tGT\t1\n")
