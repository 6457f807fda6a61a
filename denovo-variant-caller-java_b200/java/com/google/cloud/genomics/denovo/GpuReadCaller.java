/*
 * GpuReadCaller.java — drop-in for ReadCaller (ReadCaller.java:57-130) behind the reference's own plugin
 * seam: "derive from DenovoCaller, register in DenovoCallers" (README.rst:128-130).
 * SOURCE ONLY (no JDK in the build image).  Lives in package com.google.cloud.genomics.denovo next to
 * DenovoCaller because DenovoShared's getters and parseLine/writeCalls are package-private.
 *
 * The reference submits one task per candidate line; here the candidate file is read once, the pileup
 * counts of ALL accepted candidates are packed into one off-heap buffer and ONE downcall returns the
 * compacted call records, which are formatted exactly like BayesInfer.java:57-68 / ReadCaller.java:126-127.
 */
package com.google.cloud.genomics.denovo;

import com.google.cloud.genomics.denovo.DenovoUtil.Chromosome;
import com.google.cloud.genomics.denovo.DenovoUtil.Genotype;
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
import java.util.List;

public class GpuReadCaller extends DenovoCaller {

  /** Supplies the packed uint16 [n][3][4] pileup counts (DAD, MOM, CHILD x A,C,G,T) of the candidates. */
  public interface PileupSource {
    void fill(List<CallHolder> candidates, MemorySegment countsOut, DenovoB200 engine);
  }

  private final DenovoShared shared;
  private final DenovoB200 engine;
  private final PileupSource source;

  public GpuReadCaller(DenovoShared shared, String libraryPath, PileupSource source) {
    this.shared = shared;
    this.source = source;
    this.engine = new DenovoB200(libraryPath, shared.getSequenceErrorRate(), shared.getDenovoMutationRate(),
        shared.getLrtThreshold(), shared.getInferMethod().ordinal(), 0);
  }

  @Override
  public void execute() throws ParseException, IOException {
    shared.getLogger().info("---- Starting Bayesian Read Caller ----");
    final File inputFile = DenovoUtil.getNormalizedFile(shared.getInputFileName());
    final File outputFile = DenovoUtil.getNormalizedFile(shared.getOutputFileName());
    List<CallHolder> candidates = new ArrayList<>();
    int lineCount = 0;
    try (BufferedReader inputReader = new BufferedReader(new FileReader(inputFile));
        PrintWriter outputWriter = new PrintWriter(outputFile)) {
      for (String line; (line = inputReader.readLine()) != null;) {
        CallHolder callHolder = parseLine(line);
        if (lineCount % DenovoUtil.READ_LOG_FREQ == 0 && lineCount > 0) {
          shared.getLogger().info(String.format("%d Read candidates processed", lineCount));
        }
        lineCount++;
        if (!shared.getChromosomes().contains(Chromosome.valueOf(callHolder.chromosome.toUpperCase()))) {
          continue;
        }
        candidates.add(callHolder);
      }
      long n = candidates.size();
      if (n > 0) {
        try (Arena arena = Arena.ofConfined()) {
          MemorySegment counts = arena.allocate(n * 24, 16);
          source.fill(candidates, counts, engine);
          MemorySegment calls = arena.allocate(n * DenovoB200.CALL.byteSize(), 8);
          long nCalls = engine.inferCounts(counts, n, calls, n);
          StringBuilder sb = new StringBuilder();
          for (long i = 0; i < nCalls; i++) {
            MemorySegment c = calls.asSlice(i * DenovoB200.CALL.byteSize(), DenovoB200.CALL.byteSize());
            CallHolder h = candidates.get((int) c.get(ValueLayout.JAVA_LONG, 0));
            sb.append(String.format("%s,%d,%s%n", h.chromosome, h.position, details(c)));
          }
          writeCalls(outputWriter, sb.toString());
        }
      }
    } finally {
      engine.close();
    }
    shared.getLogger().info("---- Read caller terminated ----");
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
