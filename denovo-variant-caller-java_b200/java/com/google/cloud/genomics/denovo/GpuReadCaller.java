/*
 * GpuReadCaller.java — drop-in for ReadCaller (ReadCaller.java:40-221) behind the reference's own plugin seam:
 * "derive from DenovoCaller, register in DenovoCallers" (README.rst:128-130, DenovoCallers.java:43-45).
 * SOURCE ONLY (no JDK in the build image); the compiled, tested twin is the C++ GpuReadCaller of
 * host/denovo_host.hpp.  Lives in package com.google.cloud.genomics.denovo next to DenovoCaller because
 * DenovoShared's getters and parseLine / writeCalls are package-private.
 *
 * The reference submits one task per candidate line, each doing three read searches, three ReadSummary
 * constructions and one BayesInfer.infer (ReadCaller.java:111-130).  Here the candidate file is read once and
 *   1. the reads at ALL accepted candidates are packed into off-heap buffers (getReadMap, :155-186, stays the
 *      reference's own API client);
 *   2. ONE downcall builds every ReadSummary on the GPU (dnv_pileup = `new ReadSummary(reads, position)` x3 per
 *      candidate, ReadSummary.java:45-59);
 *   3. ONE downcall runs BayesInfer.infer for all candidates, interval-sharded over the devices of a dnv_group,
 *      the call records gathered in input order (dnv_group_infer_counts);
 * and the returned records are formatted exactly like BayesInfer.java:57-68 / ReadCaller.java:126-127.
 */
package com.google.cloud.genomics.denovo;

import com.google.api.services.genomics.model.Read;
import com.google.cloud.genomics.denovo.DenovoUtil.Chromosome;
import com.google.cloud.genomics.denovo.DenovoUtil.Genotype;
import com.google.cloud.genomics.denovo.DenovoUtil.TrioMember;
import com.google.cloud.genomics.denovo.gpu.DenovoB200;

import java.io.BufferedReader;
import java.io.File;
import java.io.FileReader;
import java.io.IOException;
import java.io.PrintWriter;
import java.lang.foreign.Arena;
import java.lang.foreign.MemorySegment;
import java.lang.foreign.ValueLayout;
import java.text.ParseException;
import java.util.ArrayList;
import java.util.Comparator;
import java.util.List;
import java.util.Map;

public class GpuReadCaller extends DenovoCaller {

  /** Where the reads at a candidate come from: the reference's API client (ReadCaller.getReadMap) by default. */
  public interface ReadSource {
    Map<TrioMember, List<Read>> readsAt(String chromosome, Long position) throws IOException;
  }

  private final DenovoShared shared;
  private final String libraryPath;
  private final int[] devices;
  private final ReadSource source;

  public GpuReadCaller(DenovoShared shared, String libraryPath, int[] devices, ReadSource source) {
    this.shared = shared;
    this.libraryPath = libraryPath;
    this.devices = devices;
    this.source = source;
  }

  /** One aligned read, reduced to what ReadSummary looks at (ReadSummary.java:45-59). */
  private static final class PackedRead {
    final long start; final String bases; final int person; // packed person order 0 DAD, 1 MOM, 2 CHILD
    PackedRead(long start, String bases, int person) { this.start = start; this.bases = bases; this.person = person; }
  }

  private static int packedPerson(TrioMember m) {
    return m == TrioMember.DAD ? 0 : m == TrioMember.MOM ? 1 : 2;
  }

  @Override
  public void execute() throws ParseException, IOException {
    shared.getLogger().info("---- Starting Bayesian Read Caller ----");
    final File inputFile = DenovoUtil.getNormalizedFile(shared.getInputFileName());
    final File outputFile = DenovoUtil.getNormalizedFile(shared.getOutputFileName());
    List<CallHolder> candidates = new ArrayList<>();
    int lineCount = 0;
    try (BufferedReader inputReader = new BufferedReader(new FileReader(inputFile));
        PrintWriter outputWriter = new PrintWriter(outputFile);
        DenovoB200 engine = new DenovoB200(libraryPath, shared.getSequenceErrorRate(), shared.getDenovoMutationRate(),
            shared.getLrtThreshold(), shared.getInferMethod().ordinal(), devices)) {
      for (String line; (line = inputReader.readLine()) != null;) {
        CallHolder callHolder = parseLine(line);
        if (lineCount % DenovoUtil.READ_LOG_FREQ == 0 && lineCount > 0) {
          shared.getLogger().info(String.format("%d Read candidates processed", lineCount));
        }
        lineCount++;
        if (!shared.getChromosomes().contains(Chromosome.valueOf(callHolder.chromosome.toUpperCase()))) {
          continue; // ReadCaller.java:86-89
        }
        candidates.add(callHolder);
      }
      final int n = candidates.size();
      if (n > 0) {
        try (Arena arena = Arena.ofConfined()) {
          MemorySegment counts = engine.allocPinned((long) n * 24);         // uint16 [n][DAD,MOM,CHILD][A,C,G,T]
          MemorySegment calls = engine.allocPinned((long) n * DenovoB200.CALL.byteSize());
          try {
            pileupAll(engine, arena, candidates, counts);
            long nCalls = engine.groupInferCounts(counts, n, calls, n);
            StringBuilder sb = new StringBuilder();
            for (long i = 0; i < nCalls; i++) {
              MemorySegment c = calls.asSlice(i * DenovoB200.CALL.byteSize(), DenovoB200.CALL.byteSize());
              CallHolder h = candidates.get((int) c.get(ValueLayout.JAVA_LONG, 0));
              String details = details(c);
              shared.getLogger().fine(String.format("%s,%d,%s", h.chromosome, h.position, details));
              sb.append(String.format("%s,%d,%s%n", h.chromosome, h.position, details)); // ReadCaller.java:126-127
            }
            writeCalls(outputWriter, sb.toString());
          } finally {
            engine.freePinned(counts);
            engine.freePinned(calls);
          }
        }
      }
    }
    shared.getLogger().info("---- Read caller terminated ----");
  }

  /**
   * getReadMap + getReadSummaryMap (ReadCaller.java:113-116,138-146) for ALL candidates in one pileup call.
   * The reference builds each ReadSummary from the reads of THAT candidate's own search only, so every candidate
   * gets its own stretch of a virtual axis: candidate k sits at (k + 1) * STRIDE and its reads keep their offset
   * to it.  Reads of different candidates can then never reach each other's site, a read returned by two
   * searches is counted once per search — exactly as in the reference — and the sites are ascending by
   * construction, as dnv_pileup wants them.
   */
  private static final long STRIDE = 1L << 20; // far above any read length

  private void pileupAll(DenovoB200 engine, Arena arena, List<CallHolder> candidates, MemorySegment countsOut)
      throws IOException {
    final int n = candidates.size();
    List<PackedRead> reads = new ArrayList<>();
    for (int k = 0; k < n; k++) {
      CallHolder h = candidates.get(k);
      long site = (k + 1) * STRIDE;
      for (Map.Entry<TrioMember, List<Read>> e : source.readsAt(h.chromosome, h.position).entrySet()) {
        for (Read r : e.getValue()) {
          long offset = h.position - r.getAlignment().getPosition().getPosition(); // ReadSummary.java:48
          reads.add(new PackedRead(site - offset, r.getAlignedSequence(), packedPerson(e.getKey())));
        }
      }
    }
    reads.sort(Comparator.<PackedRead>comparingInt(r -> r.person).thenComparingLong(r -> r.start));
    long nReads = reads.size(), nBases = 0;
    for (PackedRead r : reads) nBases += r.bases.length();
    MemorySegment start = arena.allocate(Math.max(nReads, 1) * 8, 8), off = arena.allocate((nReads + 1) * 8, 8);
    MemorySegment bases = arena.allocate(Math.max(nBases, 1), 16), person = arena.allocate(Math.max(nReads, 1), 1);
    long b = 0;
    for (int k = 0; k < nReads; k++) {
      PackedRead r = reads.get(k);
      start.setAtIndex(ValueLayout.JAVA_LONG, k, r.start);
      off.setAtIndex(ValueLayout.JAVA_LONG, k, b);
      person.set(ValueLayout.JAVA_BYTE, k, (byte) r.person);
      MemorySegment.copy(r.bases.getBytes(java.nio.charset.StandardCharsets.US_ASCII), 0, bases, ValueLayout.JAVA_BYTE, b,
          r.bases.length());
      b += r.bases.length();
    }
    off.setAtIndex(ValueLayout.JAVA_LONG, nReads, b);
    MemorySegment sitePos = arena.allocate((long) n * 8, 8);
    for (int k = 0; k < n; k++) sitePos.setAtIndex(ValueLayout.JAVA_LONG, k, (k + 1) * STRIDE);
    engine.pileup(start, off, bases, person, nReads, sitePos, n, countsOut); // IllegalArgumentException on a base like 'N'
  }

  /** readCounts=CHILD:{..};MOM:{..};DAD:{..},maxGenoType=[D, M, C],isDenovo=true (BayesInfer.java:57-68) */
  private static String details(MemorySegment call) {
    String[] who = {"DAD", "MOM", "CHILD"};
    String[] maps = new String[3];
    for (int p = 0; p < 3; p++) {
      StringBuilder sb = new StringBuilder("{");
      for (int a = 0; a < 4; a++) {
        int n = Short.toUnsignedInt(call.get(ValueLayout.JAVA_SHORT, 8 + 2L * (p * 4 + a)));
        if (n > 0) {
          if (sb.length() > 1) sb.append(", ");
          sb.append("ACGT".charAt(a)).append('=').append(n);
        }
      }
      maps[p] = who[p] + ":" + sb.append('}');
    }
    Genotype[] g = Genotype.values();
    return String.format("readCounts=%s;%s;%s,maxGenoType=[%s, %s, %s],isDenovo=%b", maps[2], maps[1], maps[0],
        g[call.get(ValueLayout.JAVA_BYTE, 32)], g[call.get(ValueLayout.JAVA_BYTE, 33)],
        g[call.get(ValueLayout.JAVA_BYTE, 34)], true);
  }
}
