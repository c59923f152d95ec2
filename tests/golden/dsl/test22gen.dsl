-- Test for a clustered index select followed by a second predicate
--
-- Query in SQL:
-- SELECT sum(col1) FROM tbl4 WHERE (col3 >= 18 and col3 < 218) AND (col2 >= 392 and col2 < 2392);
--
s1=select(db1.tbl4.col3,18,218)
f1=fetch(db1.tbl4.col2,s1)
s2=select(s1,f1,392,2392)
f2=fetch(db1.tbl4.col1,s2)
print(f2)
a1=sum(f2)
print(a1)
